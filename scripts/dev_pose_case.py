"""Dev aid: runs one pose-fitting problem (npz with c2, c3, K) on the GPU library and on the oracle with the same seed
and prints the local-optimisation traces side by side.  python scripts/dev_pose_case.py case.npz SEED [max_iters]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from epos_b200 import _lib, posefit  # noqa: E402
from oracle import posefit as opf  # noqa: E402

d = np.load(sys.argv[1])
seed = int(sys.argv[2])
mi = int(sys.argv[3]) if len(sys.argv) > 3 else 400
c2, c3, K = d['c2'], d['c3'], d['K']
n = c2.shape[0]
dev = torch.device('cuda:0')
p = posefit.default_params(max_iters=mi)
f = posefit.PoseFitter(dev, 1, p)
poses, lab = f.fit(torch.from_numpy(c2).to(dev), torch.from_numpy(c3).to(dev), torch.zeros(1, dtype=torch.int32, device=dev),
                   torch.tensor([n], dtype=torch.int32, device=dev), torch.from_numpy(K.reshape(1, 3, 3)).to(dev),
                   torch.tensor([seed], dtype=torch.int64, device=dev))
torch.cuda.synchronize()
rec = poses.cpu().numpy()[0]
gl = lab.cpu().numpy()
tr = np.zeros((16, 72), np.int32)
_lib.check(_lib.lib().epos_fit_debug_trace(f._ws_ptr, 1, 0, tr.ctypes.data), 'trace')
op, ol, _, st = opf.find6DPoses(c2, c3, K, max_model_number=1, seed=seed, return_stats=True, threshold=4.0,
                                min_triangle_area=0.0, max_iters=mi)
otr = np.zeros((16, 72), np.int32)
rounds = opf.lib().ora_last_trace(otr.ctypes.data_as(C.POINTER(C.c_int)), 16)
print('gpu rec', rec[12:], 'oracle', st, int(ol.sum()), 'label diff', int((gl != ol).sum()))
for r in range(max(rounds, int(rec[15]))):
    print('round', r + 1)
    print('  gpu   ', tr[r, :5].tolist(), tr[r, 5:65].reshape(20, 3).tolist())
    print('  oracle', otr[r, :5].tolist(), otr[r, 5:65].reshape(20, 3).tolist())
