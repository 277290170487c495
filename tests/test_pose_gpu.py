"""GPU parity tests of correspondence extraction and pose fitting against the oracle, through the C ABI.

Bars (north_star / SURVEY.md 8d): correspondences bit-exact (indices, f64 coordinates, f32 confidences, order);
poses: identical inlier sets, identical iteration / graph-cut counts, (R, t) within 1e-4 (relative for t) of the
oracle under a shared seed and shared (deterministic) neighbour graph."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DEV = 'cuda:0'


def _softmax(x, axis=-1):
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return (e / e.sum(axis=axis, keepdims=True)).astype(np.float32)


def _random_maps(rng, B, h, w, O, F, sharp=2.0):
    oc = _softmax(rng.standard_normal((B, h, w, O + 1)) * sharp)
    fc = _softmax(rng.standard_normal((B, h, w, O, F)) * sharp)
    fl = rng.standard_normal((B, h, w, O, F, 3)).astype(np.float32)
    return oc, fc, fl


def _check_corresp(oc, fc, fl, store, O, F, max_corr, min_obj_conf=0.1, min_rel=0.5):
    from epos_b200 import corresp
    from oracle import corresp as ocorr
    B, h, w = oc.shape[:3]
    cap = h * w * F if not max_corr else max_corr
    ex = corresp.CorrespExtractor(DEV, O, F, store, cap=cap, max_correspondences=max_corr, min_obj_conf=min_obj_conf,
                                  min_frag_rel_conf=min_rel)
    bc = ex(torch.from_numpy(oc).to(DEV), torch.from_numpy(fc).to(DEV), torch.from_numpy(fl).to(DEV))
    torch.cuda.synchronize()
    counts, totals = bc.counts.cpu().numpy(), bc.totals.cpu().numpy()
    ids = store.dp_model['obj_ids']
    J = len(ids)
    nrows = 0
    for b in range(B):
        ref = ocorr.establish_many_to_many(oc[b], fc[b], fl[b], ids, ids, store.frag_centers, store.frag_sizes, 0.25,
                                           min_obj_conf, min_rel, only_annotated_objs=False)
        for j, oid in enumerate(ids):
            s = b * J + j
            r = ref.get(oid)
            if r is None:
                assert counts[s] == 0 and totals[s] == 0
                continue
            assert totals[s] == r['coord_2d'].shape[0]
            r = ocorr.select_top_k(r, max_corr if max_corr else None)
            n = r['coord_2d'].shape[0]
            assert counts[s] == n, (b, oid, counts[s], n)
            nrows += n
            assert np.array_equal(bc.px[s, :n].cpu().numpy(), r['pixel'])
            assert np.array_equal(bc.frag[s, :n].cpu().numpy(), r['frag_id'])
            assert np.array_equal(bc.coord_2d[s, :n].cpu().numpy(), r['coord_2d'])
            assert np.array_equal(bc.coord_3d[s, :n].cpu().numpy(), r['coord_3d'])
            assert np.array_equal(bc.conf[s, :n].cpu().numpy(), r['conf'])
            assert np.array_equal(bc.conf_obj[s, :n].cpu().numpy(), r['conf_obj'])
            assert np.array_equal(bc.conf_frag[s, :n].cpu().numpy(), r['conf_frag'])
    return nrows


@pytest.mark.parametrize('O,F,h,w,max_corr', [(3, 64, 120, 160, 0), (3, 64, 120, 160, 700), (2, 8, 33, 47, 0),
                                              (5, 256, 30, 40, 512), (1, 64, 60, 80, 4096), (4, 40, 31, 37, 100)])
def test_corresp_random_maps(O, F, h, w, max_corr):
    from epos_b200 import synthetic
    rng = np.random.default_rng(O * 100 + F)
    store = synthetic.model_store(O, F)
    oc, fc, fl = _random_maps(rng, 2, h, w, O, F)
    n = _check_corresp(oc, fc, fl, store, O, F, max_corr, min_obj_conf=0.3 if O < 3 else 0.1)
    assert n > 0


def test_corresp_planted_maps_with_ties_and_empty_objects():
    """Constant confidences: every row ties, so top-K is decided by the descending-index rule alone."""
    from epos_b200 import synthetic
    O, F = 4, 64
    store = synthetic.model_store(O, F)
    oc, fc, fl, gt = synthetic.planted_maps(2, O, F, store, synthetic.default_K(), seed=5, objs_per_image=2)
    assert _check_corresp(oc, fc, fl, store, O, F, 0) > 0
    assert _check_corresp(oc, fc, fl, store, O, F, 300) > 0
    assert _check_corresp(oc, fc, fl, store, O, F, 4096) > 0


@pytest.mark.parametrize('O,F,h,w,max_corr', [(3, 64, 60, 80, 0), (3, 64, 60, 80, 500), (2, 256, 30, 40, 1024)])
def test_corresp_lazy_localisation_head(O, F, h, w, max_corr):
    """epos_corresp_lazy_loc never sees pred_frag_loc: it evaluates the 1x1 localisation head (model.py:448-456) at the
    surviving rows from the decoder features.  Against epos_corresp on the MATERIALISED head (features x weights in f64):
    row sets, order, 2D coordinates and confidences bit-identical; 3D coordinates to fp32 rounding of the dot product."""
    from epos_b200 import corresp, synthetic
    rng = np.random.default_rng(7 + F)
    B, C = 2, 256
    store = synthetic.model_store(O, F)
    oc, fc, _ = _random_maps(rng, B, h, w, O, F)
    feat = torch.from_numpy(rng.standard_normal((B * h * w, C)).astype(np.float32)).to(DEV)
    hi = feat.to(torch.bfloat16)
    lo = (feat - hi.float()).to(torch.bfloat16)
    fs = torch.stack([hi, lo]).contiguous()                              # split-bf16 [2, M, C]
    w_loc = torch.from_numpy((rng.standard_normal((O * F * 3, C)) * 0.05).astype(np.float32)).to(DEV)
    b_loc = torch.from_numpy(rng.standard_normal(O * F * 3).astype(np.float32)).to(DEV)
    a64 = hi.double() + lo.double()
    fl = (a64 @ w_loc.double().T + b_loc.double()).float().view(B, h, w, O, F, 3).contiguous()
    cap = h * w * F if not max_corr else max_corr
    kw = dict(cap=cap, max_correspondences=max_corr, min_obj_conf=0.2)
    ocd, fcd = torch.from_numpy(oc).to(DEV), torch.from_numpy(fc).to(DEV)
    ref = corresp.CorrespExtractor(DEV, O, F, store, **kw)(ocd, fcd, fl)
    got = corresp.CorrespExtractor(DEV, O, F, store, **kw)(ocd, fcd, None, lazy_loc=(fs, w_loc, b_loc))
    torch.cuda.synchronize()
    counts = ref.counts.cpu().numpy()
    assert np.array_equal(counts, got.counts.cpu().numpy()) and np.array_equal(ref.totals.cpu().numpy(), got.totals.cpu().numpy())
    assert counts.sum() > 1000
    for s_, n in enumerate(counts):
        for k in ('px', 'frag', 'coord_2d', 'conf', 'conf_obj', 'conf_frag'):
            assert torch.equal(getattr(ref, k)[s_, :n], getattr(got, k)[s_, :n]), (s_, k)
        a, b = ref.coord_3d[s_, :n], got.coord_3d[s_, :n]
        if n:
            assert float((a - b).abs().max()) < 1e-4 * max(1.0, float(a.abs().max()))     # mm; fp32 dot of 256 terms
    # without a bias
    got2 = corresp.CorrespExtractor(DEV, O, F, store, **kw)(ocd, fcd, None, lazy_loc=(fs, w_loc, None))
    fl2 = (a64 @ w_loc.double().T).float().view(B, h, w, O, F, 3).contiguous()
    ref2 = corresp.CorrespExtractor(DEV, O, F, store, **kw)(ocd, fcd, fl2)
    n0 = int(counts[0])
    assert float((ref2.coord_3d[0, :n0] - got2.coord_3d[0, :n0]).abs().max()) < 1e-4 * float(ref2.coord_3d[0, :n0].abs().max())


def test_establish_many_to_many_dropin():
    from epos_b200 import corresp, synthetic
    from oracle import corresp as ocorr
    rng = np.random.default_rng(3)
    O, F = 3, 16
    store = synthetic.model_store(O, F)
    oc, fc, fl = _random_maps(rng, 1, 24, 32, O, F)
    got = corresp.establish_many_to_many(torch.from_numpy(oc[0]).to(DEV), torch.from_numpy(fc[0]).to(DEV),
                                         torch.from_numpy(fl[0]).to(DEV), [1, 3], store, 0.25, 0.1, 0.5,
                                         only_annotated_objs=True)
    ref = ocorr.establish_many_to_many(oc[0], fc[0], fl[0], [1, 3], store.dp_model['obj_ids'], store.frag_centers,
                                       store.frag_sizes, 0.25, 0.1, 0.5, only_annotated_objs=True)
    assert sorted(got) == sorted(ref)
    for oid in ref:
        for k in ('px_id', 'frag_id', 'coord_2d', 'coord_3d', 'conf', 'conf_obj', 'conf_frag'):
            assert np.array_equal(got[oid][k].cpu().numpy(), ref[oid][k]), (oid, k)


# ---------------------------------------------------------------------------------------------------
def _fit_both(x2d, x3d, K, seed, **kw):
    from epos_b200 import posefit
    from oracle import posefit as opf
    gp, gl, gs = posefit.find6DPoses(x2d, x3d, K, max_model_number=1, seed=seed, **kw)
    rec = posefit.find6DPoses.last_record.copy()
    op, ol, os_, st = opf.find6DPoses(x2d, x3d, K, max_model_number=1, seed=seed, return_stats=True, **kw)
    return gp, gl, rec, op, ol, st


def _assert_same_fit(gp, gl, rec, op, ol, st, tol=1e-4):
    assert gp.shape == op.shape, (gp.shape, op.shape, rec, st)
    assert int(rec[13]) == st['iterations'], (rec[13], st)
    assert int(rec[15]) == st['graph_cuts'], (rec[15], st)
    assert np.array_equal(gl, ol), int((gl != ol).sum())
    if op.shape[0] == 3:
        assert np.abs(gp[:, :3] - op[:, :3]).max() < tol
        assert np.linalg.norm(gp[:, 3] - op[:, 3]) < tol * max(1.0, np.linalg.norm(op[:, 3]))


def test_fit_golden_fixtures_match_oracle():
    for name, seeds in (('pnp16.json', (0, 1, 2)), ('pose6dscene.json', (0, 1)), ('tless.json', (0, 1, 2, 3))):
        g = json.load(open(os.path.join(GOLDEN, name)))
        c, K = np.array(g['corrs']), np.array(g['K'])
        for seed in seeds:
            out = _fit_both(c[:, :2], c[:, 2:], K, seed, threshold=4.0, min_triangle_area=0.0)
            _assert_same_fit(*out)
        if name == 'pnp16.json':
            np.testing.assert_allclose(out[0][:, :3], np.array(g['R_gcransac']), atol=1e-5)
            np.testing.assert_allclose(out[0][:, 3], np.array(g['t_gcransac']), atol=1e-3)


def test_fit_planted_scene_matches_oracle_and_ground_truth():
    from epos_b200 import synthetic
    from oracle import corresp as ocorr
    O, F = 3, 64
    store = synthetic.model_store(O, F)
    K = synthetic.default_K()
    oc, fc, fl, gt = synthetic.planted_maps(2, O, F, store, K, seed=7)
    ids = store.dp_model['obj_ids']
    checked = 0
    for b in range(2):
        ref = ocorr.establish_many_to_many(oc[b], fc[b], fl[b], ids, ids, store.frag_centers, store.frag_sizes, 0.25,
                                           0.1, 0.5, only_annotated_objs=False)
        for oid, d in ref.items():
            d = ocorr.select_top_k(d, 4096)
            out = _fit_both(d['coord_2d'], d['coord_3d'], K, seed=100 * b + oid, threshold=4.0, min_triangle_area=0.0)
            _assert_same_fit(*out)
            R, t = gt[b][oid]
            assert np.abs(out[0][:, :3] - R).max() < 2e-2 and np.linalg.norm(out[0][:, 3] - t) < 5.0
            checked += 1
    assert checked >= 4


def test_fit_config5_iteration_budget_matches_oracle():
    """BASELINE configs[4]: 2000 RANSAC iterations.  Planted scene (consensus: coverage exit and LO path), the T-LESS
    fixture and a no-consensus set (all 2000 iterations, 25 chunks of passes): identical inliers / iterations / cuts."""
    from epos_b200 import synthetic
    from oracle import corresp as ocorr
    O, F = 2, 64
    store = synthetic.model_store(O, F)
    K = synthetic.default_K()
    oc, fc, fl, gt = synthetic.planted_maps(1, O, F, store, K, seed=17)
    ids = store.dp_model['obj_ids']
    ref = ocorr.establish_many_to_many(oc[0], fc[0], fl[0], ids, ids, store.frag_centers, store.frag_sizes, 0.25, 0.1, 0.5,
                                       only_annotated_objs=False)
    checked = 0
    for oid, d in ref.items():
        d = ocorr.select_top_k(d, 4096)
        out = _fit_both(d['coord_2d'], d['coord_3d'], K, seed=oid, threshold=4.0, min_triangle_area=0.0, max_iters=2000)
        _assert_same_fit(*out)
        checked += 1
    assert checked >= 1
    g = json.load(open(os.path.join(GOLDEN, 'tless.json')))
    c, Kt = np.array(g['corrs']), np.array(g['K'])
    for seed in (0, 1):
        out = _fit_both(c[:, :2], c[:, 2:], Kt, seed, threshold=4.0, min_triangle_area=0.0, max_iters=2000)
        _assert_same_fit(*out)
    rng = np.random.default_rng(12)
    x2d = 4.0 * (rng.integers(0, 160, (1500, 2)) + 0.5)
    x3d = rng.uniform(-100, 100, (1500, 3))
    out = _fit_both(x2d, x3d, K, 5, threshold=4.0, min_triangle_area=0.0, max_iters=2000)
    assert out[5]['iterations'] >= 2000
    _assert_same_fit(*out)


def test_fit_without_consensus_runs_all_iterations():
    rng = np.random.default_rng(11)
    K = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]])
    for n, seed in ((200, 1), (900, 2), (64, 3)):
        x2d = 4.0 * (rng.integers(0, 160, (n, 2)) + 0.5)
        x3d = rng.uniform(-100, 100, (n, 3))
        out = _fit_both(x2d, x3d, K, seed, threshold=4.0, min_triangle_area=0.0)
        assert out[5]['iterations'] >= 400
        _assert_same_fit(*out)


def test_fit_small_and_degenerate_inputs():
    from epos_b200 import posefit
    K = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]])
    rng = np.random.default_rng(2)
    # fewer than 6 correspondences: no pose (scripts/infer.py:417-422)
    p, lab, s = posefit.find6DPoses(rng.uniform(0, 600, (5, 2)), rng.uniform(-50, 50, (5, 3)), K, max_model_number=1)
    assert p.shape == (0, 4) and lab.sum() == 0 and s.shape == (0,)
    # all points identical: every sample is degenerate, nothing is found
    x2d = np.tile([[100.0, 100.0]], (50, 1))
    x3d = np.tile([[1.0, 2.0, 3.0]], (50, 1))
    out = _fit_both(x2d, x3d, K, 0, threshold=4.0, min_triangle_area=0.0)
    assert out[0].shape == (0, 4) and out[3].shape == (0, 4)


@pytest.mark.parametrize('O,F,B,per_image,min_found', [(4, 64, 3, 3, 6), (21, 64, 8, 5, 30)])
def test_batch_fitter_matches_oracle_pipeline(O, F, B, per_image, min_found):
    """BASELINE configs[2]-shaped post-processing: maps -> correspondences -> poses for a batch, against the oracle;
    the second case is the benched shape (8 images x 21 object slots = 168 problems in one launch)."""
    from epos_b200 import posefit, synthetic
    from oracle import pipeline
    store = synthetic.model_store(O, F)
    K = synthetic.default_K()
    oc, fc, fl, gt = synthetic.planted_maps(B, O, F, store, K, seed=21, objs_per_image=per_image)
    bf = posefit.BatchFitter(DEV, O, F, store, K, max_correspondences=2048, seed=9)
    recs = bf.fit_maps(torch.from_numpy(oc).to(DEV), torch.from_numpy(fc).to(DEV), torch.from_numpy(fl).to(DEV))
    torch.cuda.synchronize()
    recs = recs.cpu().numpy()
    pp = pipeline.PostProcess(O, F, seed=9, model_store=store, K=K, max_correspondences=2048)
    found = 0
    for b in range(B):
        corr = pp.corresp({'pred_obj_conf': oc, 'pred_frag_conf': fc, 'pred_frag_loc': fl}, b)
        ref = pp.fit(corr, image_index=b, images_per_batch=B, batch_index=0)
        for j, oid in enumerate(store.dp_model['obj_ids']):
            g, r = recs[b, j], ref[oid]
            assert g[14] == r[14] and g[12] == r[12] and g[13] == r[13] and g[15] == r[15], (b, oid, g[12:], r[12:])
            if r[14] == 1.0:
                found += 1
                assert np.abs(g[:12] - r[:12]).max() < 1e-4 * max(1.0, np.abs(r[:12]).max())
                if oid in gt[b]:
                    R, t = gt[b][oid]
                    assert np.abs(g[:12].reshape(3, 4)[:, :3] - R).max() < 2e-2
    assert found >= min_found


@pytest.mark.parametrize('F', [16, 128])
def test_engine_full_path_postprocessing_matches_oracle_on_its_own_maps(F):
    """Engine = model.predict -> corresp -> fit on the device; the oracle post-processes the SAME maps (copied to the
    host), so the comparison is exact although the CNN itself only matches the f32 oracle to 1e-3."""
    from epos_b200 import engine, model, synthetic, weights as W
    from oracle import pipeline
    O, B = 3, 2
    w = W.random_init(O, F, seed=2, bn='random', logits_std=0.5)
    store = synthetic.model_store(O, F)
    K = synthetic.default_K()
    eng = engine.Engine(w, O, F, DEV, stages=engine.STAGES_FULL, model_store=store, K=K, max_correspondences=1024, seed=4,
                        lazy_loc=False)
    img = torch.from_numpy(W.synthetic_images(B, seed=6, height=160, width=224)).to(DEV)
    out = eng.run_device(img)
    torch.cuda.synchronize()
    recs = out['poses'].cpu().numpy()
    assert recs.shape == (B, O, 16)
    maps = {k: out[k].cpu().numpy() for k in (model.PRED_OBJ_CONF, model.PRED_FRAG_CONF, model.PRED_FRAG_LOC)}
    pp = pipeline.PostProcess(O, F, seed=4, model_store=store, K=K, max_correspondences=1024)
    nprob = 0
    for b in range(B):
        ref = pp.fit(pp.corresp(maps, b), image_index=b, images_per_batch=B, batch_index=0)
        for j, oid in enumerate(store.dp_model['obj_ids']):
            g, r = recs[b, j], ref[oid]
            assert g[13] == r[13] and g[14] == r[14] and g[12] == r[12] and g[15] == r[15], (b, oid, g[12:], r[12:])
            nprob += r[13] > 0
            if r[14] == 1.0:
                assert np.abs(g[:12] - r[:12]).max() < 1e-4 * max(1.0, np.abs(r[:12]).max())
    assert nprob >= 2
    # second batch uses the next stream keys
    out2 = eng.run_device(img)
    assert eng._fitter.batch_index == 2


def test_lazy_engine_matches_materialising_engine_and_oracle_fit():
    """The engine's default path never materialises pred_frag_loc.  Its correspondences must be those of the
    materialising engine (rows / order / confidences identical, 3D coordinates to fp32 rounding), and its pose records
    must equal the oracle's fit of the engine's OWN correspondences (identical inliers / iterations / graph cuts)."""
    from epos_b200 import engine, model, synthetic, weights as W
    from oracle import posefit as opf
    O, F, B = 3, 64, 2
    w = W.random_init(O, F, seed=2, bn='random', logits_std=0.5)
    store = synthetic.model_store(O, F)
    K = synthetic.default_K()
    img = torch.from_numpy(W.synthetic_images(B, seed=6, height=160, width=224)).to(DEV)
    res = {}
    for lazy in (False, True):
        eng = engine.Engine(w, O, F, DEV, stages=engine.STAGES_FULL, model_store=store, K=K, max_correspondences=1024, seed=4,
                            lazy_loc=lazy)
        out = eng.run_device(img)
        torch.cuda.synchronize()
        assert (model.PRED_FRAG_LOC in out) == (not lazy) and (model.LAZY_FRAG_LOC in out) == lazy
        bc = eng._fitter.corr
        res[lazy] = (out['poses'].cpu().numpy(), bc.counts.cpu().numpy(), bc.px.cpu().numpy(), bc.frag.cpu().numpy(),
                     bc.conf.cpu().numpy(), bc.coord_2d.cpu().numpy(), bc.coord_3d.cpu().numpy())
    a, b = res[False], res[True]
    assert np.array_equal(a[1], b[1]) and a[1].sum() > 100
    nprob = 0
    for s_, n in enumerate(a[1]):
        for k in (2, 3, 4, 5):
            assert np.array_equal(a[k][s_, :n], b[k][s_, :n])
        if n:
            assert np.abs(a[6][s_, :n] - b[6][s_, :n]).max() < 1e-4 * max(1.0, np.abs(a[6][s_, :n]).max())
        if n >= 6:
            bi, j = divmod(s_, O)
            op, ol, _, st = opf.find6DPoses(b[5][s_, :n], b[6][s_, :n], K, max_model_number=1, seed=(4 << 32) + bi * O + j,
                                            return_stats=True, threshold=4.0, min_triangle_area=0.0)
            g = b[0][bi, j]
            assert int(g[13]) == st['iterations'] and int(g[15]) == st['graph_cuts'] and int(g[14]) == st['found']
            if st['found']:
                assert int(g[12]) == int(ol.sum())
                assert np.abs(g[:12].reshape(3, 4) - op).max() < 1e-4 * max(1.0, np.abs(op).max())
            nprob += 1
    assert nprob >= 2


def test_pipelined_engine_equals_serial_engine():
    """pipelined=True overlaps pose fitting of batch i (side stream) with the CNN of batch i+1; every record must be
    bit-identical to the serial engine's, for every batch, through run_device and through run_host/flush."""
    from epos_b200 import engine, synthetic, weights as W
    O, F, B = 3, 16, 2
    w = W.random_init(O, F, seed=2, bn='random', logits_std=0.5)
    store = synthetic.model_store(O, F)
    K = synthetic.default_K()
    imgs = [torch.from_numpy(W.synthetic_images(B, seed=20 + i, height=160, width=224)) for i in range(4)]
    res = {}
    for mode in (False, True):
        eng = engine.Engine(w, O, F, DEV, stages=engine.STAGES_FULL, model_store=store, K=K, max_correspondences=1024,
                            seed=4, pipelined=mode)
        assert eng.pipelined == mode
        outs = [eng.run_device(im.to(DEV)) for im in imgs]
        eng.join()
        torch.cuda.synchronize()
        res[mode] = [o['poses'].cpu().numpy().copy() for o in outs]
        host = []
        for im in imgs:
            r = eng.run_host(im.pin_memory())
            if r is not None:
                host.append(r.numpy().copy())
        last = eng.flush()
        if last is not None:
            host.append(last.numpy().copy())
        assert len(host) == len(imgs)
        res[(mode, 'host')] = host
    for a, b in zip(res[False], res[True]):
        assert a.shape == (B, O, 16) and np.array_equal(a, b)
    assert sum(float(r[..., 14].sum()) for r in res[False]) >= 1          # something was actually fitted
    # run_host sees batches 4..7 of each engine's seed stream: same keys in both engines
    for a, b in zip(res[(False, 'host')], res[(True, 'host')]):
        assert np.array_equal(a, b)


def _engine_records(w, O, F, store, K, imgs, Ks, pipelined, graphs):
    """Pose records of every batch through run_device (read one batch at a time) and then through run_host / flush."""
    from epos_b200 import engine
    eng = engine.Engine(w, O, F, DEV, stages=engine.STAGES_FULL, model_store=store, K=K, max_correspondences=1024,
                        seed=4, pipelined=pipelined, graphs=graphs)
    recs = []
    for im, k in zip(imgs, Ks):
        o = eng.run_device(im.to(DEV), K=k)
        eng.join()
        torch.cuda.synchronize()
        recs.append(o['poses'].cpu().numpy().copy())
    host = []
    for im, k in zip(imgs, Ks):
        r = eng.run_host(im.pin_memory(), K=k)
        if r is not None:
            host.append(r.numpy().copy())
    last = eng.flush() if pipelined else None
    if last is not None:
        host.append(last.numpy().copy())
    assert len(host) == len(imgs)
    torch.cuda.synchronize()
    return recs, host, eng


@pytest.mark.parametrize('pipelined', [False, True])
def test_graph_engine_equals_eager_engine(pipelined):
    """graphs=True replays the per-batch pipeline from CUDA graphs (two buffer sets, device-side stream-key counter).
    Every pose record must be bit-identical to the serial eager engine's, batch after batch, through run_device and
    run_host, back to back without host synchronisation, and a change of the intrinsics between batches must take effect."""
    from epos_b200 import engine, synthetic, weights as W
    O, F, B = 3, 16, 2
    w = W.random_init(O, F, seed=2, bn='random', logits_std=0.5)
    store = synthetic.model_store(O, F)
    K = synthetic.default_K()
    K2 = K.copy(); K2[0, 0] *= 1.1; K2[1, 1] *= 1.1
    imgs = [torch.from_numpy(W.synthetic_images(B, seed=20 + i, height=160, width=224)) for i in range(5)]
    Ks = [None, None, K2, K2, None]
    ref_dev, ref_host, _ = _engine_records(w, O, F, store, K, imgs, Ks, pipelined=False, graphs=False)
    assert sum(float(r[..., 14].sum()) for r in ref_dev) >= 1
    assert not np.array_equal(ref_dev[1], ref_dev[2])                  # K2 changes the poses
    for graphs in ((False, True) if pipelined else (True,)):
        dev, host, eng = _engine_records(w, O, F, store, K, imgs, Ks, pipelined=pipelined, graphs=graphs)
        for i, (a, b) in enumerate(zip(ref_dev, dev)):
            assert a.shape == (B, O, 16) and np.array_equal(a, b), ('run_device', 'graphs' if graphs else 'eager', i)
        for i, (a, b) in enumerate(zip(ref_host, host)):
            assert np.array_equal(a, b), ('run_host', 'graphs' if graphs else 'eager', i, np.abs(a - b).max())
        if graphs:
            assert eng.launch_count() > 0 and eng.graph_launches > 100
        # back-to-back run_device calls (no host synchronisation in between): the last record is still the right one
        eng2 = engine.Engine(w, O, F, DEV, stages=engine.STAGES_FULL, model_store=store, K=K, max_correspondences=1024,
                             seed=4, pipelined=pipelined, graphs=graphs)
        outs = [eng2.run_device(im.to(DEV), K=k) for im, k in zip(imgs, Ks)]
        eng2.join()
        torch.cuda.synchronize()
        assert np.array_equal(outs[-1]['poses'].cpu().numpy(), ref_dev[-1]), ('back to back', graphs)


def test_full_size_pipelined_graph_engine_is_deterministic_under_overlap():
    """The benchmark configuration (8 x 480 x 640, 21 objects x 64 fragments, CUDA graphs, pose fitting of batch i under
    the CNN of batch i + 1): the same batch fed 10 times back to back must give bit-identical head maps every time --
    the CTA-pair GEMMs take their tiles from a queue while the fit kernel's CTAs come and go on the same SMs -- and the
    records of a second engine fed the same sequence must be identical (same seeds per batch index)."""
    from epos_b200 import engine, model, synthetic, weights as W
    O, F, B = 21, 64, 8
    w = W.random_init(O, F, seed=0, logits_std=300.0)
    store, K = synthetic.model_store(O, F), synthetic.default_K()
    img = torch.from_numpy(W.synthetic_images(B, seed=3)).to(DEV)
    runs = []
    for rep in range(2):
        eng = engine.Engine(w, O, F, DEV, stages=engine.STAGES_FULL, model_store=store, K=K, max_correspondences=4096,
                            seed=99, pipelined=True, graphs=True)
        maps, recs = [], []
        outs = [eng.run_device(img) for _ in range(10)]          # back to back: fit of batch i overlaps CNN of batch i + 1
        eng.join()
        torch.cuda.synchronize()
        # the two buffer sets hold the maps of the last two batches; every batch's records were cloned
        for o in outs[-2:]:
            maps.append((o[model.PRED_OBJ_CONF].clone(), o[model.PRED_FRAG_CONF].clone()))
        assert torch.equal(maps[0][0], maps[1][0]) and torch.equal(maps[0][1], maps[1][1])
        eng2_first = eng.run_device(img)                         # one more, alone (nothing overlapping its CNN)
        eng.join()
        torch.cuda.synchronize()
        assert torch.equal(eng2_first[model.PRED_OBJ_CONF], maps[0][0]) and torch.equal(eng2_first[model.PRED_FRAG_CONF], maps[0][1])
        runs.append([o['poses'].cpu().numpy().copy() for o in outs])
    for i, (a, b) in enumerate(zip(*runs)):
        assert np.array_equal(a, b), i
    assert sum(float(r[..., 14].sum()) for r in runs[0]) > 0


def test_graph_engine_cnn_only():
    from epos_b200 import engine, model, weights as W
    O, F, B = 2, 8, 2
    w = W.random_init(O, F, seed=5, bn='random', logits_std=0.5)
    imgs = [torch.from_numpy(W.synthetic_images(B, seed=30 + i, height=96, width=128)).to(DEV) for i in range(3)]
    ref = engine.Engine(w, O, F, DEV, stages=engine.STAGES_CNN)
    eng = engine.Engine(w, O, F, DEV, stages=engine.STAGES_CNN, graphs=True)
    for im in imgs:
        a = ref.run_device(im)
        b = eng.run_device(im)
        torch.cuda.synchronize()
        for k in (model.PRED_OBJ_CONF, model.PRED_FRAG_CONF, model.PRED_FRAG_LOC, model.PRED_OBJ_LABEL):
            assert torch.equal(a[k], b[k]), k


# ---------------------------------------------------------------------------------------------------
# multi-instance fitting (Progressive-X + PEARL)
def _multi_both(x2d, x3d, K, seed, mm, **kw):
    from epos_b200 import posefit
    from oracle import posefit as opf
    gp, gl, gs = posefit.find6DPoses(x2d, x3d, K, max_model_number=mm, max_model_number_for_optimization=5, seed=seed, **kw)
    state = posefit.find6DPoses.last_multi_state.copy()
    op, ol, os_, st = opf.find6DPoses(x2d, x3d, K, max_model_number=mm, max_model_number_for_optimization=5, seed=seed,
                                      return_stats=True, **kw)
    return gp, gl, gs, state, op, ol, os_, st


def _assert_same_multi(gp, gl, gs, state, op, ol, os_, st, tol=1e-4):
    """Same instances (count, order), same labeling, poses within 1e-4 (relative for t), same number of proposals and
    RANSAC iterations.  PEARL iteration counts may differ by the refits that change a converged model in the last bits."""
    assert gp.shape == op.shape, (gp.shape, op.shape, state, st)
    assert int(state[0]) == st['proposals'] and int(state[1]) == st['accepted'] and int(state[2]) == st['ransac_iterations'], (state, st)
    assert np.array_equal(gl, ol), (int((gl != ol).sum()), np.bincount(gl).tolist(), np.bincount(ol).tolist())
    for k in range(op.shape[0] // 3):
        a, b = gp[3 * k:3 * k + 3], op[3 * k:3 * k + 3]
        assert np.abs(a[:, :3] - b[:, :3]).max() < tol, (k, np.abs(a - b).max())
        assert np.linalg.norm(a[:, 3] - b[:, 3]) < tol * max(1.0, np.linalg.norm(b[:, 3]))
    np.testing.assert_allclose(gs, os_, rtol=1e-6, atol=1e-6)


def _two_instance_scene(rng, n_per=350, n_out=250, noise=0.8):
    from epos_b200 import synthetic
    K = synthetic.default_K()
    pts = []
    poses = []
    for _ in range(2):
        rv = rng.normal(size=3)
        rv *= rng.uniform(0.3, 2.5) / np.linalg.norm(rv)
        R = synthetic.rodrigues(rv)
        t = np.array([rng.uniform(-150, 150), rng.uniform(-100, 100), rng.uniform(600, 1100)])
        X = rng.normal(size=(n_per, 3))
        X *= 100.0 / np.linalg.norm(X, axis=1, keepdims=True)
        Xc = X @ R.T + t
        front = (Xc - t)[:, 2] < 0                                    # visible half of the sphere
        X, Xc = X[front], Xc[front]
        uv = (Xc[:, :2] / Xc[:, 2:3]) * np.array([K[0, 0], K[1, 1]]) + np.array([K[0, 2], K[1, 2]])
        uv = 4.0 * (np.floor(uv / 4.0) + 0.5) + rng.normal(0, noise, uv.shape) * 0     # pixel-centre grid like corresp.py
        pts.append(np.concatenate([uv + rng.normal(0, noise, uv.shape), X], 1))
        poses.append((R, t))
    out = np.concatenate([4.0 * (rng.integers(0, 160, (n_out, 2)) + 0.5), rng.uniform(-100, 100, (n_out, 3))], 1)
    allp = np.concatenate(pts + [out])
    return allp[rng.permutation(len(allp))], K, poses


def test_multi_instance_tless_fixture_matches_oracle():
    g = json.load(open(os.path.join(GOLDEN, 'tless.json')))
    c, K = np.array(g['corrs']), np.array(g['K'])
    for mm, seeds in ((2, (0, 1)), (3, (0,))):
        for seed in seeds:
            out = _multi_both(c[:, :2], c[:, 2:], K, seed, mm, threshold=4.0, min_triangle_area=0.0)
            assert out[4].shape[0] == 3 * mm
            _assert_same_multi(*out)


def test_multi_instance_planted_scenes_match_oracle_and_ground_truth():
    rng = np.random.default_rng(5)
    for trial in range(3):
        pts, K, poses = _two_instance_scene(rng)
        out = _multi_both(pts[:, :2], pts[:, 2:], K, 10 + trial, 2, threshold=4.0, min_triangle_area=0.0)
        _assert_same_multi(*out)
        gp = out[0]
        assert gp.shape[0] == 6
        for R, t in poses:
            errs = [np.abs(gp[3 * k:3 * k + 3, :3] - R).max() for k in range(2)]
            assert min(errs) < 3e-2, errs
    # a bound above the number of real instances: the extra proposals are rejected (Tanimoto / too few inliers)
    pts, K, poses = _two_instance_scene(rng, n_out=100)
    out = _multi_both(pts[:, :2], pts[:, 2:], K, 3, 4, threshold=4.0, min_triangle_area=0.0)
    _assert_same_multi(*out)


def test_multi_instance_no_consensus_and_argument_checks():
    from epos_b200 import posefit
    K = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]])
    rng = np.random.default_rng(4)
    x2d, x3d = rng.uniform(0, 640, (40, 2)), rng.uniform(-50, 50, (40, 3))
    out = _multi_both(x2d, x3d, K, 0, 2, threshold=0.02, min_triangle_area=0.0, max_iters=50)
    assert out[0].shape == (0, 4) and out[4].shape == (0, 4) and int(out[3][0]) == 101
    _assert_same_multi(*out)
    with pytest.raises(ValueError):
        posefit.find6DPoses(x2d, x3d, K, max_model_number=0)


def test_multi_instance_sequential_fitting_matches_oracle():
    """More instances than max_model_number_for_optimization, or -1 (DETECTION): spedUpFitting (progressive_x.h:265-391):
    proposals on the points the earlier ones left, neighbourhood rebuilt over all 7 columns, no PEARL."""
    from epos_b200 import posefit
    from oracle import posefit as opf
    g = json.load(open(os.path.join(GOLDEN, 'tless.json')))
    c, K = np.array(g['corrs']), np.array(g['K'])
    rng = np.random.default_rng(8)
    pts, K2, poses = _two_instance_scene(rng)
    for (x2d, x3d, Kc, mm, mopt, seed) in ((c[:, :2], c[:, 2:], K, 4, 3, 0), (pts[:, :2], pts[:, 2:], K2, 7, 5, 1),
                                           (pts[:, :2], pts[:, 2:], K2, -1, 5, 2)):
        kw = dict(threshold=4.0, min_triangle_area=0.0, max_model_number=mm, max_model_number_for_optimization=mopt, seed=seed)
        gp, gl, gs = posefit.find6DPoses(x2d, x3d, Kc, **kw)
        state = posefit.find6DPoses.last_multi_state.copy()
        op, ol, os_, st = opf.find6DPoses(x2d, x3d, Kc, return_stats=True, **kw)
        assert st['sped_up'] == 1 and gp.shape == op.shape and op.shape[0] >= 6, (gp.shape, op.shape, state, st)
        assert int(state[0]) == st['proposals'] and int(state[2]) == st['ransac_iterations'], (state, st)
        assert not gl.any() and not ol.any() and not gs.any() and not os_.any()
        for k in range(op.shape[0] // 3):
            a, b = gp[3 * k:3 * k + 3], op[3 * k:3 * k + 3]
            assert np.abs(a[:, :3] - b[:, :3]).max() < 1e-4, (k, np.abs(a - b).max())
            assert np.linalg.norm(a[:, 3] - b[:, 3]) < 1e-4 * max(1.0, np.linalg.norm(b[:, 3]))
