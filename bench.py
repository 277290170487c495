#!/usr/bin/env python
"""bench.py -- images/sec of the EPOS per-image inference hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cnn|full]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of `--batch` synthetic 640x480 images PER GPU (weak
scaling: images are sharded across ranks, no data-path collective except the all-gather of pose records).
Prints ONE JSON line (rank 0).  Keys follow the driver contract; see DESIGN.md "Measurement".

  value      images/s, inputs already resident in HBM (CUDA events, max over ranks)
  e2e        images/s through Engine.run_host(): pinned host images -> H2D -> hot path -> D2H of the result
  roofline   the dominant kernel (tcgen05 pointwise GEMM): algorithmic FLOPs / CUDA-event time of its launches
  cpu_baseline  the oracle (CPU restatement of the reference path) on the host cores, bounded sample

--impl reference times the reference's CPU path (oracle port: TF-1.12 / OpenCV / Eigen are not installable
here) on the host cores for the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 480, 640
TRUNK_FLOP = {'xception_65': 401.97e9,          # trunk + ASPP + decoder, per image
              'resnet_v1_50_beta': 277.70e9}
# Random-init logit weights (reference: truncated_normal(0.01), model.py:437) give uniform heads -> obj_conf = 1/22 < tau_a
# -> ZERO correspondences.  The full workload therefore scales the logit initialiser so that the
# random features produce a varied, mostly no-consensus correspondence load (0 .. >30k rows per object, top-K 4096):
# every object runs all 400 RANSAC iterations and its graph-cut budget, i.e. the expensive case for pose fitting.
HEAD_STD_FULL = 300.0
MAX_CORR = 4096


def head_std(kind, backbone):
    """Logit initialiser stddev of the full workload (None = the reference's 0.01 for CNN-only runs)."""
    if kind != 'full':
        return None
    return HEAD_STD_FULL


def head_flop(O, F):
    return 2.0 * 120 * 160 * 256 * ((O + 1) + O * F + 3 * O * F)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=None, choices=['cnn', 'full'])
    ap.add_argument('--batch', type=int, default=8, help='images per GPU per step')
    ap.add_argument('--objs', type=int, default=21)
    ap.add_argument('--frags', type=int, default=64)
    ap.add_argument('--backbone', default='xception_65', choices=['xception_65', 'resnet_v1_50_beta'],
                    help='model_variant (BASELINE configs[3] uses resnet_v1_50_beta)')
    ap.add_argument('--max-iters', type=int, default=400, help='RANSAC iteration budget (configs[4]: 2000)')
    ap.add_argument('--cpu-images', type=int, default=3, help='images in the cpu_baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--serial', action='store_true', help='no CNN / pose-fitting overlap (single stream)')
    ap.add_argument('--no-graphs', dest='graphs', action='store_false', default=True,
                    help='launch every kernel from Python instead of replaying the captured CUDA graphs')
    ap.add_argument('--secondary', dest='secondary', action='store_true', default=None,
                    help='also time BASELINE configs[3] (ResNet-50-beta, 8/GPU) and configs[4] (30x256, 2000 iterations, '
                         '16/GPU) for a few steps each (default: on when no workload flag is given)')
    ap.add_argument('--no-secondary', dest='secondary', action='store_false')
    return ap.parse_args()


def default_workload():
    try:
        from epos_b200 import posefit  # noqa: F401
        return 'full'
    except Exception:
        return 'cnn'


def workload_name(kind, B, O, F, backbone='xception_65'):
    if (O, F) == (30, 256):
        return ('BASELINE configs[4] per-GPU shape: batch=%d/GPU 640x480 synthetic RGB, random-init %s, 30 objects / 256 '
                'fragments, %s' % (B, backbone, 'CNN-only forward' if kind == 'cnn' else
                                   'full CNN + corresp + GC-RANSAC pose fitting'))
    if backbone != 'xception_65':
        return ('BASELINE configs[3]-shaped: batch=%d/GPU 640x480 synthetic RGB, random-init %s backbone, %d-object / '
                '%d-fragment heads, %s' % (B, backbone, O, F, 'CNN-only forward' if kind == 'cnn' else
                                           'full CNN + corresp + GC-RANSAC pose fitting'))
    if kind == 'cnn':
        return ('BASELINE configs[1]: batch=%d/GPU 640x480 synthetic RGB, random-init Xception-65 f64 '
                '(%d objects x %d fragments heads), CNN-only forward (model.predict)' % (B, O, F))
    return ('BASELINE configs[2]: batch=%d/GPU 640x480 synthetic RGB, random-init Xception-65 f64, %d-object / '
            '%d-fragment heads, full CNN + corresp + GC-RANSAC pose fitting' % (B, O, F))


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, reasons, mx, pw = [], set(), None, []
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx = float(c[2]); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


# ------------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle): cpu_baseline leg and --impl reference
# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(kind, O, F, n_images, threads, seed=0, backbone='xception_65'):
    """Times the oracle on `n_images` images, one at a time like scripts/infer.py (batch 1); the first image is
    dropped as warm-up like infer.py:741-749.  Returns (images_per_s, per-stage seconds)."""
    import numpy as np
    import torch
    from epos_b200 import weights as Wt
    from oracle import cnn as ocnn
    torch.set_num_threads(threads)
    w = Wt.random_init(O, F, seed=seed, logits_std=head_std(kind, backbone), model_variant=backbone)
    net = ocnn.Oracle(w, model_variant=backbone)
    fit = None
    if kind == 'full':
        from oracle import pipeline as opipe
        fit = opipe.PostProcess(O, F, seed=seed, max_correspondences=MAX_CORR)
    stages = {'prediction': 0.0, 'establish_corr': 0.0, 'fitting': 0.0}
    timed = 0
    for i in range(n_images + 1):
        img = Wt.synthetic_images(1, seed=seed + 100 + i)
        t0 = time.perf_counter()
        out = net.predict(img, O, F)
        t1 = time.perf_counter()
        t2 = t3 = t1
        if fit is not None:
            corr = fit.corresp(out)
            t2 = time.perf_counter()
            fit.fit(corr, image_index=i)
            t3 = time.perf_counter()
        if i == 0:
            continue
        timed += 1
        stages['prediction'] += t1 - t0
        stages['establish_corr'] += t2 - t1
        stages['fitting'] += t3 - t2
    total = sum(stages.values())
    return timed / total, {k: v / timed for k, v in stages.items()}


def run_reference(args, kind):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(10, cores)            # scripts/infer.py:695-698 pins TF to 10 intra/inter-op threads
    n = max(1, args.steps)
    # warm-up images are untimed; each "step" of the reference is ONE image (infer.py is batch 1).
    ips, stages = cpu_reference_run(kind, args.objs, args.frags, n, threads, backbone=args.backbone)
    line = {
        'impl': 'reference', 'metric': 'images/sec (640x480, Xception-65 f64 + PnP-RANSAC)', 'value': ips,
        'unit': 'images/s', 'n_gpus': args.gpus, 'steps': n, 'warmup': 1, 'ms_per_step': 1e3 / ips,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(kind, args.batch, args.objs, args.frags, args.backbone),
                   'note': 'reference CPU path = oracle port (TF-1.12/OpenCV-3.4/Eigen not installable); '
                           'batch 1 per step as in scripts/infer.py:610'},
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d images, one per step, first extra image dropped as warm-up; per-stage s/img %s'
                                   % (n, json.dumps({k: round(v, 4) for k, v in stages.items()}))},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
class Cfg:
    """One measured workload (a BASELINE.json config at its per-GPU shape)."""

    def __init__(self, kind, batch, objs, frags, backbone, max_iters, steps, warmup):
        self.kind, self.batch, self.objs, self.frags = kind, batch, objs, frags
        self.backbone, self.max_iters, self.steps, self.warmup = backbone, max_iters, steps, warmup


def measure(cfg, ctx, serial=False, want_roofline=True, graphs=True):
    """Builds the engine for `cfg`, times `value` (device-resident inputs) and `e2e` (host buffers) over exactly
    cfg.steps steps after cfg.warmup warm-up steps, max over ranks; rank 0 also measures the rooflines of the two
    dominant kernels (tcgen05 GEMM: tensor; fit_kernel: HBM on algorithmic bytes).  Returns a dict (rank 0) / None."""
    import torch
    import torch.distributed as dist
    from epos_b200 import _lib, engine, weights as Wt
    from epos_b200 import dist as edist
    from epos_b200 import model as emodel
    world, rank, local, dev, lib, side_group = (ctx[k] for k in ('world', 'rank', 'local', 'dev', 'lib', 'side_group'))
    kind, B, O, F = cfg.kind, cfg.batch, cfg.objs, cfg.frags

    # weights: generated on rank 0 and broadcast once over NCCL
    w = edist.broadcast_weights(Wt.random_init(O, F, seed=0, logits_std=head_std(kind, cfg.backbone),
                                               model_variant=cfg.backbone)
                                if rank == 0 else None, O, F, dev, world, rank, model_variant=cfg.backbone)
    store = K = None
    if kind == 'full':
        from epos_b200 import synthetic
        store = synthetic.model_store(O, F)
        K = synthetic.default_K()
    opts = emodel.ModelOptions(Wt.head_channels(O, F), model_variant=cfg.backbone)
    fit_params = None
    if kind == 'full' and cfg.max_iters != 400:
        from epos_b200 import posefit
        fit_params = posefit.default_params()
        fit_params.max_iters = cfg.max_iters
    eng = engine.Engine(w, O, F, dev, stages=engine.STAGES_FULL if kind == 'full' else engine.STAGES_CNN,
                        model_store=store, K=K, seed=1234 + rank, max_correspondences=MAX_CORR, model_options=opts,
                        fit_params=fit_params, pipelined=not serial, graphs=graphs,
                        post_fit=(lambda poses: edist.all_gather_poses(poses, world, group=side_group)) if world > 1 else None)
    del w

    # inputs: NROT distinct batches (> L2 in total) rotated between steps, both pinned-host and device copies
    NROT = 5
    host_batches = [torch.from_numpy(Wt.synthetic_images(B, seed=1000 * rank + i)).pin_memory() for i in range(NROT)]
    dev_batches = [h.to(dev) for h in host_batches]
    torch.cuda.synchronize()

    def barrier():
        # drain BOTH streams first: the pose all-gathers run on the engine's side stream, and NCCL operations of one rank
        # must not be in flight on two streams in an order that can differ between ranks
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ("value") ----
    for i in range(cfg.warmup):
        out = eng.run_device(dev_batches[i % NROT])
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(cfg.steps):
        out = eng.run_device(dev_batches[i % NROT])
    eng.join()                         # the timed region ends when the last batch's pose records (and all-gather) are done
    e1.record()
    barrier()
    launches = int(eng.launch_count() - l0)          # kernels replayed through CUDA graphs included
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * cfg.steps / (ms * 1e-3)
    torch.cuda.synchronize()

    # ---- end-to-end through the public host API ("e2e") ----
    res0 = eng.result_tensor(out)
    res_pinned = torch.empty(res0.shape, dtype=res0.dtype).pin_memory()
    res_pinned2 = torch.empty(res0.shape, dtype=res0.dtype).pin_memory()     # pipelined: batch i lands while i+1 runs
    for i in range(cfg.warmup):
        eng.run_host(host_batches[i % NROT], res_pinned if i % 2 == 0 else res_pinned2)
    eng.flush()
    barrier()
    e0.record()
    for i in range(cfg.steps):
        eng.run_host(host_batches[i % NROT], res_pinned if i % 2 == 0 else res_pinned2)
    eng.flush()                        # host holds the last batch's result
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e = {'value': world * B * cfg.steps / (e2e_ms * 1e-3), 'unit': 'images/s',
           'h2d_bytes_per_step': int(host_batches[0].numel() * host_batches[0].element_size()),
           'd2h_bytes_per_step': int(res_pinned.numel() * res_pinned.element_size()),
           'ms_per_step': e2e_ms / cfg.steps}

    # ---- rooflines of the dominant kernels: per-launch CUDA events, engine serial (each kernel alone on the GPU) ----
    roof = roof_ransac = None
    if rank == 0 and want_roofline:
        import ctypes as C
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
        except Exception:
            pass
        eng.net.gemm_events = []
        nat = max(1, min(cfg.steps, 5))
        was_pipelined, eng.pipelined = eng.pipelined, False       # kernel timed alone: no pose-fitting CTAs beside it
        was_graphs, eng.graphs = eng.graphs, False                # per-launch events need eager launches
        post_fit, eng.post_fit = eng.post_fit, None               # rank 0 only: no collective in this pass
        fit_ms = fit_bytes = prep_ms = 0.0
        if kind == 'full':
            _lib.check(lib.epos_fit_enable_timing(1), 'epos_fit_enable_timing')
        for i in range(nat):
            eng.run_device(dev_batches[i % NROT])
            if kind == 'full':
                a, b = C.c_float(0), C.c_float(0)
                _lib.check(lib.epos_fit_last_kernel_ms(C.byref(a), C.byref(b)), 'epos_fit_last_kernel_ms')
                prep_ms += a.value
                fit_ms += b.value
                fit_bytes += eng._fitter._fitter.algorithmic_bytes(B * eng._fitter.J)
        torch.cuda.synchronize()
        if kind == 'full':
            _lib.check(lib.epos_fit_enable_timing(0), 'epos_fit_enable_timing')
        eng.pipelined, eng.post_fit, eng.graphs = was_pipelined, post_fit, was_graphs
        evs, eng.net.gemm_events = eng.net.gemm_events, None
        gemm_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in evs)
        gemm_flop = sum(2.0 * m * n * k for _, _, m, n, k in evs)
        peak = peaks.get('bf16_tflops_sustained') or 1400.0
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'gemm_traffic.json')) as f:
                tj = json.load(f)
            traffic = {'dram_bytes_per_launch': tj['bytes_per_launch'], 'algorithmic_mb_per_launch': tj['algorithmic_mb_per_launch'],
                       'source': tj['source']}
        except Exception:
            pass
        achieved = gemm_flop / (gemm_ms * 1e-3) / 1e12
        roof = {'bound': 'tensor', 'kernel': 'pw_gemm_kernel (tcgen05 split-bf16 pointwise conv, %d launches/step)'
                % (len(evs) // nat), 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peaks else 'fallback',
                'traffic': traffic,
                'mma_issue_frac': 3.0 * achieved / peak,
                'note': 'achieved = algorithmic fp32-equivalent FLOPs (2MNK per launch, SURVEY 8d) / summed per-launch '
                        'event time; each product costs 3 bf16 MMAs (error-compensated split), mma_issue_frac = 3x',
                'gemm_ms_per_step': gemm_ms / nat, 'share_of_step': (gemm_ms / nat) / (ms / cfg.steps)}
        if kind == 'full' and fit_ms > 0:
            hbm = peaks.get('hbm_gbs') or 6500.0
            ach = fit_bytes / (fit_ms * 1e-3) / 1e9
            traffic_r = None
            try:
                with open(os.path.join(ROOT, 'profiles', 'ransac_traffic.json')) as f:
                    traffic_r = json.load(f)
            except Exception:
                pass
            roof_ransac = {
                'bound': 'hbm', 'kernel': 'fit_kernel (persistent CTA per (image, object) problem: GC-RANSAC main loop, '
                                          'graph-cut LO, final LSQ/LM), 1 launch/step',
                'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback', 'traffic': traffic_r,
                'algorithmic_mb_per_launch': fit_bytes / nat / 1e6, 'fit_ms_per_launch': fit_ms / nat,
                'prep_ms_per_launch': prep_ms / nat,
                'note': 'SURVEY 8d: every model scored over a problem\'s N points costs N x 40 B (u_n, v_n, x, y, z f64) '
                        '+ N x 8 B (pixel id); models counted on the device (epos_fit_debug_state cols 16-18: main loop '
                        'hypotheses, LO trials, final re-scorings).  The point set is shared-memory resident, so DRAM '
                        'traffic is far below the algorithmic bytes; the kernel is FP64-issue / latency bound.'}

    res = None
    if rank == 0:
        flop_img = TRUNK_FLOP[cfg.backbone] + head_flop(O, F)
        res = {'value': value, 'ms_per_step': ms / cfg.steps, 'steps': cfg.steps, 'warmup': cfg.warmup, 'e2e': e2e,
               'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'roofline_ransac': roof_ransac,
               'model_tflops_algorithmic': value * flop_img / 1e12 / world,
               'config': {'workload': workload_name(kind, B, O, F, cfg.backbone), 'images_per_gpu_per_step': B,
                          'global_batch': B * world,
                          'l2': 'inputs rotate over %d distinct batches (%.0f MB) and per-layer activations (>=112 MB at '
                                'B=8) exceed the 126 MB L2' % (NROT, NROT * host_batches[0].numel() * 4 / 1e6),
                          'algorithmic_gflop_per_image': flop_img / 1e9,
                          'heads': 'random-init; logit initialiser stddev %s' % (head_std(kind, cfg.backbone) or 0.01),
                          'ransac_max_iters': cfg.max_iters if kind == 'full' else None,
                          'max_correspondences': MAX_CORR if kind == 'full' else None,
                          'parallelism': 'image-sharded dp%d' % world,
                          'pipeline': 'pose fitting of batch i on a side stream under the CNN of batch i+1'
                                      if eng.pipelined else 'serial',
                          'launch': 'CUDA graphs (CNN graph + post-processing graph per buffer set, replayed)'
                                    if eng.graphs else 'eager (one ctypes launch per kernel)',
                          'parity_pin': 'pose: reference golden vectors (pnp16 / pose6dscene / tless / cv2 / BK max-flow); '
                                        'CNN: Slim conv2d_same / atrous vectors only (the reference holds no network-level '
                                        'vector; TF-1.12 not installable)'}}
    del eng, dev_batches, host_batches, out
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def run_ours(args, kind):
    import torch
    import torch.distributed as dist
    from epos_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU path)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    # the per-batch all-gather of pose records runs on the engine's side stream: give it its own communicator so that it
    # never interleaves with the main-stream collectives (barrier, timing all-reduce) of the default group
    side_group = dist.new_group(backend='nccl') if world > 1 else None
    ctx = {'world': world, 'rank': rank, 'local': local, 'dev': dev, 'lib': _lib.lib(), 'side_group': side_group}
    B, O, F = args.batch, args.objs, args.frags
    primary = measure(Cfg(kind, B, O, F, args.backbone, args.max_iters, args.steps, args.warmup), ctx, serial=args.serial,
                      graphs=args.graphs)

    # ---- the other multi-GPU configs of BASELINE.json at their per-GPU shapes (short runs, same timing rules) ----
    secondary = None
    if args.secondary and kind == 'full':
        # the engine of a secondary config is built after seconds of host-only work (weight generation): ten warm-up steps
        # bring the clocks back up before the short timed run
        sw, ss = max(10, args.warmup), max(3, min(args.steps, 10))
        secondary = {}
        for name, cfg in (('configs[3]', Cfg('full', 8, 21, 64, 'resnet_v1_50_beta', 400, ss, sw)),
                          ('configs[4]', Cfg('full', 16, 30, 256, 'xception_65', 2000, ss, sw))):
            r = measure(cfg, ctx, serial=args.serial, graphs=args.graphs)
            if rank == 0:
                r['metric'], r['unit'], r['n_gpus'] = 'images/sec', 'images/s', world
                secondary[name] = r

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        threads = min(10, cores)
        ips, stages = cpu_reference_run(kind, O, F, args.cpu_images, threads, backbone=args.backbone)
        cpu = {'value': ips, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
               'sample': '%d images of the same workload, batch 1 as scripts/infer.py, +1 warm-up image dropped; s/img %s'
                         % (args.cpu_images, json.dumps({k: round(v, 4) for k, v in stages.items()}))}

    if rank == 0:
        line = {
            'metric': 'images/sec (640x480, Xception-65 f64 + PnP-RANSAC)', 'value': primary['value'], 'unit': 'images/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': primary['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (split-bf16 x3 MMA, f32 accumulate; pose f64)', 'data': 'synthetic',
            'config': primary['config'], 'clocks': primary['clocks'], 'e2e': primary['e2e'],
            'gpu_launches': primary['gpu_launches'], 'roofline': primary['roofline'],
            'roofline_ransac': primary['roofline_ransac'], 'cpu_baseline': cpu,
            'model_tflops_algorithmic': primary['model_tflops_algorithmic'], 'secondary': secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    kind = args.workload or default_workload()
    if args.secondary is None:          # the plain driver invocation measures every multi-GPU config of BASELINE.json
        args.secondary = (args.workload is None and args.batch == 8 and args.objs == 21 and args.frags == 64 and
                          args.backbone == 'xception_65' and args.max_iters == 400)
    if args.impl == 'reference':
        run_reference(args, kind)
    else:
        run_ours(args, kind)


if __name__ == '__main__':
    main()
