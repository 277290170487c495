#!/usr/bin/env python
"""profiles/gemm_traffic.json and profiles/ransac_traffic.json (the `traffic` fields of bench.py's roofline objects) from
the ncu --set full captures of a tag: DRAM bytes read + written per launch of the dominant kernels.
usage: make_traffic_json.py <tag>   (reads gpurun_out/prof_gemm_<tag>.ncu-rep, prof_ransac_<tag>.ncu-rep)"""
import csv, json, os, subprocess, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(name):
    rep = os.path.join(root, 'gpurun_out', 'prof_%s_%s.ncu-rep' % (name, tag))
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hdr, units, data = rd[0], rd[1], rd[2:]

    def mb(d, key):
        i = hdr.index(key)
        v = float(d[i])
        return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}[units[i]]
    return [{'kernel': d[hdr.index('Kernel Name')].split('(')[0], 'grid': d[hdr.index('Grid Size')],
             'duration_us': float(d[hdr.index('gpu__time_duration.sum')]) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[units[hdr.index('gpu__time_duration.sum')]],
             'dram_read_mb': round(mb(d, 'dram__bytes_read.sum'), 2), 'dram_write_mb': round(mb(d, 'dram__bytes_write.sum'), 2)} for d in data]


g = [r for r in rows('gemm')]
mid = [r for r in g if 100 < r['duration_us'] < 135]          # the 38400 x 728 x 728 middle-flow launches
sel = mid or g
per = sum(r['dram_read_mb'] + r['dram_write_mb'] for r in sel) / len(sel)
json.dump({'kernel': 'pw_gemm_kernel<256,64>', 'source': 'gpurun_out/prof_gemm_%s.ncu-rep -> profiles/prof_gemm_%s_raw.md (ncu --set full '
           '--clock-control none, launches of the 38400x728x728 middle-flow layer, bench.py --steps 1 --no-graphs --serial)' % (tag, tag),
           'launches': sel, 'algorithmic_mb_per_launch': 225.8, 'bytes_per_launch': int(per * 1e6)},
          open(os.path.join(root, 'profiles', 'gemm_traffic.json'), 'w'), indent=1)
f = rows('ransac')
per = sum(r['dram_read_mb'] + r['dram_write_mb'] for r in f) / len(f)
json.dump({'kernel': 'fit_kernel', 'source': 'gpurun_out/prof_ransac_%s.ncu-rep -> profiles/prof_ransac_%s_raw.md (ncu --set full '
           '--clock-control none, one launch = 168 problems of the config-3 workload)' % (tag, tag), 'launches': f,
           'dram_bytes_per_launch': int(per * 1e6),
           'note': 'the point set of a problem is shared-memory resident: DRAM traffic is far below the algorithmic bytes of SURVEY 8d'},
          open(os.path.join(root, 'profiles', 'ransac_traffic.json'), 'w'), indent=1)
print(open(os.path.join(root, 'profiles', 'gemm_traffic.json')).read()[:600])
print(open(os.path.join(root, 'profiles', 'ransac_traffic.json')).read()[:600])
