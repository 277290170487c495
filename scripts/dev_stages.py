"""Developer timing of the three stages of the full path (CUDA events), same workload as bench.py --workload full."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from epos_b200 import engine, synthetic, weights as W, _lib
import bench
B, O, F = 8, 21, 64
dev = torch.device('cuda:0')
w = W.random_init(O, F, seed=0, logits_std=bench.HEAD_STD_FULL)
store, K = synthetic.model_store(O, F), synthetic.default_K()
eng = engine.Engine(w, O, F, dev, stages=engine.STAGES_FULL, model_store=store, K=K, seed=1234, max_correspondences=bench.MAX_CORR)
imgs = [torch.from_numpy(W.synthetic_images(B, seed=i)).to(dev) for i in range(3)]
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(2):
    out = eng.run_device(imgs[it])
torch.cuda.synchronize()
tot = {'cnn': 0, 'corresp': 0, 'fit': 0}
n = 4
from epos_b200 import model
for it in range(n):
    e = [ev() for _ in range(4)]
    e[0].record()
    out = eng.net.predict(imgs[it % 3], lazy_loc=eng.lazy_loc)
    e[1].record()
    bf = eng._fitter
    bf._prepare(B)
    bc = bf.extract(out[model.PRED_OBJ_CONF], out[model.PRED_FRAG_CONF], out.get(model.PRED_FRAG_LOC), lazy_loc=out.get(model.LAZY_FRAG_LOC))
    e[2].record()
    seeds = bf.seeds_for(B); bf.batch_index += 1
    poses, lab = bf._fitter.fit(bc.coord_2d.view(-1, 2), bc.coord_3d.view(-1, 3), bc.offsets, bc.counts, bf._Kdev, seeds, bf._poses, bf._labeling)
    e[3].record()
    torch.cuda.synchronize()
    tot['cnn'] += e[0].elapsed_time(e[1]); tot['corresp'] += e[1].elapsed_time(e[2]); tot['fit'] += e[2].elapsed_time(e[3])
print({k: round(v / n, 3) for k, v in tot.items()}, 'ms per batch of', B)
c = bc.counts.cpu().numpy().reshape(B, O); t = bc.totals.cpu().numpy().reshape(B, O)
print('counts img0', c[0].tolist()); print('totals img0', t[0].tolist())
r = poses.cpu().numpy().reshape(B, O, 16)
print('valid', int((r[..., 14] == 1).sum()), 'of', B * O, 'iterations mean', r[..., 13].mean(), 'graph cuts mean', r[..., 15].mean())
dbg = np.zeros((B * O, 20), np.int64)
_lib.check(_lib.lib().epos_fit_debug_state(bf._fitter._ws_ptr, B * O, dbg.ctypes.data), 'dbg')
tot_c = dbg[:, 11] + dbg[:, 12] + dbg[:, 13] + dbg[:, 14]
order = np.argsort(-tot_c)[:10]
print('slowest problems: [N, used_px, iters, passes, gcuts, lo_runs, phase, best_inl] | Mcycles main(sample, score, replay, total) cut trials final fit-in-trials')
for i in order:
    print(i, dbg[i, :8].tolist(), (dbg[i, 8:16] / 1e6).round(2).tolist())
print('sum over problems (Mcycles): main %.1f cut %.1f trials %.1f final %.1f' % tuple(dbg[:, k].sum() / 1e6 for k in (11, 12, 13, 14)))
