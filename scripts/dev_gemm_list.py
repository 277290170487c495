"""Per-shape GEMM time inside one forward pass (CUDA events around every tcgen05 GEMM launch)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from epos_b200 import model, weights as W
B, O, F = 8, 21, 64
dev = torch.device('cuda:0')
net = model.EposNet(W.random_init(O, F, seed=0), O, F, dev)
img = torch.from_numpy(W.synthetic_images(B, seed=0)).to(dev)
for _ in range(2):
    net.predict(img)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for it in range(3):
    net.gemm_events = []
    net.predict(img)
    torch.cuda.synchronize()
    for e0, e1, M, N, K in net.gemm_events:
        a = agg.setdefault((M, N, K), [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1) * 1e3
tot = sum(v[1] for v in agg.values()) / 3
print('GEMM total %.2f ms per forward' % (tot / 1e3))
for (M, N, K), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    per = us / n
    print('M=%6d N=%4d K=%4d x%2d  %7.1f us each  %6.2f ms/fwd (%4.1f%%)  %5.0f TF/s mma' % (
        M, N, K, n // 3, per, us / 3 / 1e3, 100 * us / 3 / tot, 6.0 * M * N * K / per / 1e6))
