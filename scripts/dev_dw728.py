import sys, os
sys.path.insert(0, '/root/repo')
import torch
from epos_b200 import _lib
lib = _lib.lib(); dev = torch.device('cuda:0')
B, H, W, C, rate = 8, 60, 80, 728, 2
LDX = 736; LDY = 736
xs = [torch.randn(B * H * W, LDX, device=dev) for _ in range(3)]
ys = [torch.empty(2, B * H * W, LDY, dtype=torch.bfloat16, device=dev) for _ in range(3)]
w = torch.randn(9, C, device=dev); b = torch.randn(C, device=dev)
s = torch.cuda.current_stream().cuda_stream
for i in range(6):
    _lib.check(lib.epos_dwconv3x3(xs[i % 3].data_ptr(), LDX, w.data_ptr(), b.data_ptr(), None, ys[i % 3].data_ptr(), LDY, B, H, W, C, 1, rate, 1, 0, s), 'dw')
torch.cuda.synchronize()
