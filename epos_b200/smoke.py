"""smoke(): one small invocation of the hot path on cuda:0, checked against the oracle (test infrastructure;
this module, tests/ and bench.py's CPU legs are the only importers of oracle/)."""
import numpy as np
import torch


def run():
    from . import model, posefit, synthetic, weights as W
    from oracle import cnn as ocnn
    from oracle import pipeline as opipe
    assert torch.cuda.is_available(), 'smoke() needs a CUDA device'
    dev = torch.device('cuda:0')
    # 1) forward pass of a small image through every CNN kernel
    O, F = 2, 8
    w = W.random_init(O, F, seed=3, bn='random', logits_std=0.5)
    img = W.synthetic_images(1, seed=3, height=96, width=128)
    net = model.EposNet(w, O, F, dev)
    out = net.predict(torch.from_numpy(img).to(dev))
    torch.cuda.synchronize()
    ref = ocnn.predict(w, img, O, F)
    for k in (model.PRED_OBJ_CONF, model.PRED_FRAG_CONF, model.PRED_FRAG_LOC):
        a, b = out[k].cpu().numpy(), ref[k]
        err = np.abs(a - b).max() / np.abs(b).max()
        print('[smoke] %-15s rel-to-max err %.3e' % (k, err))
        assert err < 1e-3, (k, err)
    agree = (out[model.PRED_OBJ_LABEL].cpu().numpy() == ref['pred_obj_label']).mean()
    print('[smoke] label agreement %.5f' % agree)
    assert agree > 0.999
    # 2) correspondences + GC-RANSAC on planted maps with known poses
    O, F, B = 2, 64, 1
    store = synthetic.model_store(O, F)
    K = synthetic.default_K()
    oc, fc, fl, gt = synthetic.planted_maps(B, O, F, store, K, seed=4)
    bf = posefit.BatchFitter(dev, O, F, store, K, max_correspondences=2048, seed=5)
    recs = bf.fit_maps(torch.from_numpy(oc).to(dev), torch.from_numpy(fc).to(dev), torch.from_numpy(fl).to(dev))
    torch.cuda.synchronize()
    recs = recs.cpu().numpy()
    pp = opipe.PostProcess(O, F, seed=5, model_store=store, K=K, max_correspondences=2048)
    refp = pp.fit(pp.corresp({'pred_obj_conf': oc, 'pred_frag_conf': fc, 'pred_frag_loc': fl}, 0), image_index=0,
                  images_per_batch=B, batch_index=0)
    for j, oid in enumerate(store.dp_model['obj_ids']):
        g, r = recs[0, j], refp[oid]
        print('[smoke] obj %d: valid %d inliers %d (oracle %d) iterations %d (oracle %d) |dpose| %.2e' % (
            oid, g[14], g[12], r[12], g[13], r[13], np.abs(g[:12] - r[:12]).max()))
        assert g[14] == r[14] == 1.0 and g[12] == r[12] and g[13] == r[13]
        assert np.abs(g[:12] - r[:12]).max() < 1e-4 * max(1.0, np.abs(r[:12]).max())
        R, t = gt[0][oid]
        assert np.abs(g[:12].reshape(3, 4)[:, :3] - R).max() < 2e-2
    print('[smoke] OK')
