"""Microbenchmark of the depthwise kernel on the dominant layer shapes (CUDA events, rotating buffers > L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from epos_b200 import _lib
lib = _lib.lib(); dev = torch.device('cuda:0')
shapes = [(8, 60, 80, 728, 1, 2), (8, 60, 80, 2048, 1, 12), (8, 120, 160, 256, 2, 1), (8, 240, 320, 128, 1, 1), (8, 60, 80, 1536, 1, 4),
          (8, 120, 160, 304, 1, 1), (8, 60, 80, 1024, 1, 2), (8, 120, 160, 256, 1, 1)]
for (B, H, W, C, stride, rate) in shapes:
    Ho, Wo = (H, W) if stride == 1 else ((H - 1) // 2 + 1, (W - 1) // 2 + 1)
    nb = 3
    LDX = (C + 31) // 32 * 32          # the network keeps f32 activations with 128-byte aligned rows (728 -> 736)
    xs = [torch.randn(B * H * W, LDX, device=dev) for _ in range(nb)]
    LDY = (C + 15) // 16 * 16
    ys = [torch.empty(2, B * Ho * Wo, LDY, dtype=torch.bfloat16, device=dev) for _ in range(nb)]
    w = torch.randn(9, C, device=dev); b = torch.randn(C, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    def run(i):
        _lib.check(lib.epos_dwconv3x3(xs[i % nb].data_ptr(), LDX, w.data_ptr(), b.data_ptr(), None, ys[i % nb].data_ptr(), LDY, B, H, W, C, stride, rate, int(os.environ.get('RELU_IN', '1')), 0, s), 'dw')
    for i in range(3): run(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 30
    for i in range(n): run(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    bytes_ = (B * H * W * C * 4 + B * Ho * Wo * C * 4)
    print('TH %s variant %s  B%d %dx%dx%d s%d r%d: %.1f us  %.0f GB/s (alg)' % (os.environ.get('EPOS_DW_TH', 'auto'), os.environ.get('EPOS_DW_VARIANT', 'default'), B, H, W, C, stride, rate, us, bytes_ / us / 1e3))
