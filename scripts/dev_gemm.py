"""Developer A/B of the tcgen05 pointwise GEMM on the network's dominant shapes (CUDA events, L2 flushed between
launches by rotating over buffers).  EPOS_GEMM_DEBUG bits: 1 old row-per-lane epilogue, 2 skip stores,
4 skip residual loads, 8 skip all MMAs, 16 one MMA instead of three."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from epos_b200 import _lib

dev = torch.device('cuda:0')
lib = _lib.lib()
SHAPES = [  # M, N, K, residual, f32 out, split out, name
    (38400, 728, 728, False, True, False, 'middle pw'),
    (38400, 728, 728, True, True, False, 'middle pw+res'),
    (38400, 1536, 1536, False, True, False, 'exit 1536'),
    (38400, 2048, 1536, False, True, False, 'exit 2048'),
    (38400, 256, 2048, False, False, True, 'aspp'),
    (153600, 256, 304, False, True, False, 'decoder0'),
    (153600, 4032, 256, False, True, False, 'head loc'),
    (153600, 1344, 256, False, True, False, 'head conf'),
]
flags = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 6, 8, 16]
only = os.environ.get('SHAPES')
if os.environ.get('CUSTOM'):
    # CUSTOM=MxNxK[:rfs] (r = residual, f = f32 output, s = split-bf16 output; default f)
    SHAPES = []
    for c in os.environ['CUSTOM'].split(','):
        dims, _, fl = c.partition(':')
        fl = fl or 'f'
        SHAPES.append(tuple(int(v) for v in dims.split('x')) + ('r' in fl, 'f' in fl, 's' in fl, c))


def split(t):
    hi = t.to(torch.bfloat16)
    return torch.stack([hi, (t - hi.float()).to(torch.bfloat16)]).contiguous()


def run(M, N, K, res, f32, spl, nbuf=3, iters=6):
    g = torch.Generator(device='cuda').manual_seed(0)
    a32 = [torch.randn(M, K, device=dev, generator=g) for _ in range(nbuf)]
    LDA = (K + int(os.environ.get('PADA', '16')) - 1) // int(os.environ.get('PADA', '16')) * int(os.environ.get('PADA', '16'))
    a = [torch.nn.functional.pad(split(t), (0, LDA - K)).contiguous() for t in a32]
    w32 = torch.randn(N, K, device=dev, generator=g) * 0.05
    LDW = (K + 63) // 64 * 64 if os.environ.get('PADW', '1') == '1' else K
    w = torch.nn.functional.pad(split(w32), (0, LDW - K)).contiguous()
    bias = torch.randn(N, device=dev, generator=g)
    r = [torch.randn(M, N, device=dev, generator=g) for _ in range(nbuf)] if res else None
    d = [torch.empty(M, N, device=dev) for _ in range(nbuf)] if f32 else None
    ds = [torch.empty(2, M, N, dtype=torch.bfloat16, device=dev) for _ in range(nbuf)] if spl else None
    s = torch.cuda.current_stream().cuda_stream

    def call(i):
        _lib.check(lib.epos_pwconv_gemm(a[i].data_ptr(), LDA, a[i].stride(0), w.data_ptr(), LDW, bias.data_ptr(), 0,
                                        r[i].data_ptr() if res else None, N, d[i].data_ptr() if f32 else None, N,
                                        ds[i].data_ptr() if spl else None, N, ds[i].stride(0) if spl else 0,
                                        M, N, K, 1, s), 'gemm')
    out = {}
    for fl in flags:
        os.environ['EPOS_GEMM_DEBUG'] = str(fl)
        for i in range(nbuf):
            call(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(iters):
            call(it % nbuf)
        e1.record(); torch.cuda.synchronize()
        out[fl] = e0.elapsed_time(e1) / iters * 1e3
        if fl in (0, 1):
            ref = a32[0][:4096].double() @ w32.double().t() + bias.double()
            if res:
                ref = ref + r[0][:4096].double()
            ref = torch.relu(ref)
            call(0); torch.cuda.synchronize()
            got = d[0][:4096].double() if f32 else (ds[0][0][:4096].double() + ds[0][1][:4096].double())
            err = ((got - ref).abs().max() / ref.abs().max()).item()
            assert err < 2e-5, (fl, err)
    os.environ['EPOS_GEMM_DEBUG'] = '0'
    return out


for M, N, K, res, f32, spl, name in SHAPES:
    if only and name not in only.split(','):
        continue
    t = run(M, N, K, res, f32, spl)
    mma = 3 * 2.0 * M * N * K
    print('%-14s M=%6d N=%4d K=%4d | ' % (name, M, N, K) +
          '  '.join('dbg%-3d %6.1f us (%4.0f)' % (f, us, mma / us / 1e6) for f, us in t.items()), flush=True)
