"""Developer timing: forward pass at a given batch; prints ms/batch and the parity-relevant error summary."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from epos_b200 import model, weights as W, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
O = int(sys.argv[2]) if len(sys.argv) > 2 else 21
F = 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device('cuda:0')
w = W.random_init(O, F, seed=0)
net = model.EposNet(w, O, F, dev)
img = torch.from_numpy(W.synthetic_images(B, seed=0)).to(dev)
for _ in range(2):
    out = net.predict(img)
torch.cuda.synchronize()
l0 = _lib.lib().epos_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    feat, b, h, w_ = net.forward_features(img)
e1.record(); torch.cuda.synchronize()
t_feat = e0.elapsed_time(e1) / iters
e0.record()
for _ in range(iters):
    out = net.predict(img)
e1.record(); torch.cuda.synchronize()
t_all = e0.elapsed_time(e1) / iters
print('B=%d O=%d: trunk %.2f ms, trunk+heads %.2f ms -> %.1f img/s (CNN only); launches/forward %d' % (
    B, O, t_feat, t_all, B / t_all * 1e3, (_lib.lib().epos_launch_count() - l0) // (2 * iters)))
print('trunk TFLOP/s (alg) %.1f' % (B * 401.97e9 / (t_feat * 1e-3) / 1e12))
