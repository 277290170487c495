#!/usr/bin/env python
"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).  usage: summarize_profiles.py <tag>"""
import csv, collections, os, subprocess, sys, re
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, 'gpurun_out'), os.path.join(root, 'profiles')
os.makedirs(pr, exist_ok=True)
# 1) launch list -> per-kernel totals and shares
rows = []
with open(os.path.join(go, 'launches_%s.csv' % tag)) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        rows.append((re.sub(r'\(.*', '', r['Kernel Name']), r['Grid Size'], r['Block Size'], float(r['Metric Value'])))
agg = collections.OrderedDict()
for k, g, b, t in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
with open(os.path.join(pr, 'launches_%s_summary.md' % tag), 'w') as f:
    f.write('# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none` over bench.py\n\n' % tag)
    f.write('%d launches captured, %.3f ms total (cold-cache, serialised: compare shares)\n\n' % (len(rows), tot / 1e6))
    f.write('| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|\n')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('| %s | %d | %.3f | %.1f%% | %.1f |\n' % (k, n, t / 1e6, 100 * t / tot, t / n / 1e3))
with open(os.path.join(pr, 'launches_%s.csv' % tag), 'w') as f:
    f.write('kernel,grid,block,ns\n')
    for k, g, b, t in rows:
        f.write('"%s","%s","%s",%d\n' % (k, g, b, t))
# 2) full captures -> selected raw metrics
pat = re.compile(r'dram__bytes_(read|write)\.sum$|dram__bytes_(read|write)\.sum\.per_second|gpu__dram_throughput\.avg\.pct|'
                 r'sm__pipe_tensor.*cycles_active.*pct|sm__inst_executed_pipe_tensor|sm__warps_active\.avg\.pct|launch__registers_per_thread|'
                 r'gpu__time_duration\.sum|sm__throughput\.avg\.pct|launch__grid_size|launch__block_size|l1tex__t_bytes.*sum$|lts__t_bytes\.sum$|'
                 r'smsp__cycles_active\.avg|sm__cycles_elapsed\.max|launch__occupancy_limit|smsp__warp_issue_stalled.*pct|sm__pipe_fp64|sm__inst_executed\.sum$|'
                 r'l1tex__m_xbar2l1tex_read_bytes|lts__throughput\.avg\.pct|sm__mem_tensor_cycles_active|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|'
                 r'l1tex__throughput\.avg\.pct|lts__t_sector_hit_rate\.pct|smsp__inst_executed_pipe_fp64|sm__inst_executed_pipe_fp64')
for fn in sorted(os.listdir(go)):
    if fn.endswith('_%s.ncu-rep' % tag):
        out = subprocess.run(['ncu', '-i', os.path.join(go, fn), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rd = list(csv.reader(out.splitlines()))
        hdr, units, data = rd[0], rd[1], rd[2:]
        with open(os.path.join(pr, fn.replace('.ncu-rep', '_raw.md')), 'w') as f:
            f.write('# %s: `ncu --set full --clock-control none` (selected raw metrics per captured launch)\n\n' % fn)
            for d in data:
                name = d[hdr.index('Kernel Name')]
                f.write('## %s grid=%s block=%s\n\n| metric | unit | value |\n|---|---|---|\n' % (
                    re.sub(r'\(.*', '', name), d[hdr.index('Grid Size')], d[hdr.index('Block Size')]))
                for i, h in enumerate(hdr):
                    if pat.search(h):
                        f.write('| %s | %s | %s |\n' % (h, units[i], d[i]))
                f.write('\n')
print(open(os.path.join(pr, 'launches_%s_summary.md' % tag)).read())
