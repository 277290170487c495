"""Pins the pose-fitting oracle (oracle/posefit.cpp) on every golden vector the reference holds for this path
(SURVEY.md section 4 / 8c): the 16-correspondence notebook answer, the pose6dscene and T-LESS fixtures, outputs of
cv2.solvePnP(ITERATIVE)/cv2.Rodrigues recorded from the cv2 wheel, and the reference's own BK max-flow
(oracle/_ref, built from /root/reference when present) for the graph-cut labeling."""
import os

import numpy as np
import pytest

from oracle import posefit as pf


def rot_err_deg(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))


def test_rng_is_a_pure_function_and_unique_sets():
    a = pf.rng_u64(1, 0, 2, 3, 4)
    assert a == pf.rng_u64(1, 0, 2, 3, 4) and a != pf.rng_u64(1, 1, 2, 3, 4) and a != pf.rng_u64(2, 0, 2, 3, 4)
    # published test vector of the stream definition: recomputed in pure Python
    M = (1 << 64) - 1

    def mix(z):
        z = (z + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    def ref(seed, stream, a_, b_, c_):
        z = mix(seed ^ ((stream * 0xD6E8FEB86659FD93) & M))
        z = mix((z + a_) & M)
        z = mix((z + b_) & M)
        return mix((z + c_) & M)
    for args in [(0, 0, 0, 0, 0), (1234, 1, 7, 19, 3), (M, 0, 399, 99, 50)]:
        assert pf.rng_u64(*args) == ref(*args)
    for n, k in [(3, 3), (10, 3), (25, 21), (4096, 21)]:
        s = pf.unique_set(5, 1, 2, 3, n, k)
        assert len(set(s.tolist())) == k and s.min() >= 0 and s.max() < n
    assert pf.unique_set(5, 0, 0, 0, 2, 3) is None
    # uniformity (coarse)
    h = np.bincount([pf.unique_set(9, 0, p, 0, 7, 1)[0] for p in range(7000)], minlength=7)
    assert h.min() > 850 and h.max() < 1150


def test_quartic_against_numpy_roots():
    rng = np.random.default_rng(0)
    for _ in range(300):
        kind = rng.integers(0, 3)
        if kind == 0:
            roots = rng.uniform(-2, 2, 4)
            c = np.poly(roots)[::-1] * rng.uniform(0.5, 3)
        elif kind == 1:
            re, im = rng.uniform(-1, 1), rng.uniform(0.3, 1)
            roots = rng.uniform(-2, 2, 2)
            c = np.real(np.poly([roots[0], roots[1], re + 1j * im, re - 1j * im]))[::-1]
        else:
            c = rng.standard_normal(5)
        got = pf.quartic(c)
        ref = np.roots(c[::-1])
        ref = np.sort(ref[np.abs(ref.imag) < 1e-9 * (1 + np.abs(ref.real))].real)
        assert len(got) == len(ref), (c, got, ref)
        np.testing.assert_allclose(got, ref, rtol=1e-7, atol=1e-9)


def _random_scene(rng, n, noise=0.0):
    import math
    X = rng.uniform(-80, 80, (n, 3))
    rv = rng.normal(size=3)
    rv *= rng.uniform(0.1, 2.5) / np.linalg.norm(rv)
    R, _ = pf.rodrigues_to_matrix(rv)
    t = np.array([rng.uniform(-100, 100), rng.uniform(-100, 100), rng.uniform(500, 1200)])
    K = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]])
    Y = X @ R.T + t
    uv = (Y[:, :2] / Y[:, 2:]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]]
    uv += rng.normal(0, noise, uv.shape) if noise else 0
    return X, uv, K, R, t


def test_p3p_recovers_pose_and_reprojects_sample():
    rng = np.random.default_rng(1)
    hits = 0
    for _ in range(200):
        X, uv, K, R, t = _random_scene(rng, 3)
        pts = pf.points7(uv, X, K)
        models = pf.p3p(pts, [0, 1, 2])
        assert 1 <= len(models) <= 4
        best = 1e9
        for m in models:
            Y = X @ m[:, :3].T + m[:, 3]
            np.testing.assert_allclose(Y[:, :2] / Y[:, 2:], pts[:, :2], atol=1e-7)     # exact on the sample
            assert abs(np.linalg.det(m[:, :3]) - 1) < 1e-6 and m[2, 3] >= 0
            best = min(best, np.abs(m[:, :3] - R).max() + np.abs(m[:, 3] - t).max() / 1000)
        hits += best < 1e-5
    assert hits >= 195        # the true pose is among the solutions (a few configurations are ill-conditioned)


def test_p3p_degenerate_inputs():
    K = np.eye(3)
    X = np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2.0]])                  # collinear world points
    uv = np.array([[0, 0], [0.1, 0], [0, 0.1]])
    assert len(pf.p3p(pf.points7(uv, X, K), [0, 1, 2])) == 0


def test_rodrigues_golden(golden):
    g = golden('cv2_solvepnp.json')
    for c in g['rodrigues']:
        R, J = pf.rodrigues_to_matrix(c['rvec'])
        np.testing.assert_allclose(R, np.array(c['R']), atol=1e-12)
        np.testing.assert_allclose(J, np.array(c['J']), atol=1e-9)
        back = pf.matrix_to_rodrigues(np.array(c['R']))
        np.testing.assert_allclose(back, np.array(c['back']), atol=1e-7)


def test_solvepnp_iterative_golden(golden):
    """cv2.solvePnP(SOLVEPNP_ITERATIVE) outputs recorded from the cv2 wheel (tests/golden/make_golden.py)."""
    g = golden('cv2_solvepnp.json')
    for c in g['cases']:
        X, uv = np.array(c['X']), np.array(c['uv'])
        ok, r, t = pf.solvepnp(X, uv)
        if len(X) < 6 and not c['planar']:
            continue
        assert ok == c['ok']
        Rg = np.array(c['R'])
        R, _ = pf.rodrigues_to_matrix(r)
        assert rot_err_deg(R, Rg) < 1e-4, (len(X), c['planar'])
        np.testing.assert_allclose(t, c['tvec'], rtol=1e-5, atol=1e-3)
        ok2, r2, t2 = pf.solvepnp(X, uv, guess=(c['guess_rvec'], c['guess_tvec']))
        R2, _ = pf.rodrigues_to_matrix(r2)
        Rg2, _ = pf.rodrigues_to_matrix(c['rvec_guess'])
        assert ok2 and rot_err_deg(R2, Rg2) < 1e-4
        np.testing.assert_allclose(t2, c['tvec_guess'], rtol=1e-5, atol=1e-3)


def test_pnp16_known_answer(golden):
    """example_pnp.ipynb:111-126,212-217: 16 correspondences, recorded (R, t) of find6DPoseEPOS."""
    g = golden('pnp16.json')
    c, K = np.array(g['corrs']), np.array(g['K'])
    for seed in range(5):
        poses, labels, scores = pf.find6DPoses(c[:, :2], c[:, 2:], K, threshold=g['threshold_px'],
                                               min_triangle_area=0.0, seed=seed)
        assert poses.shape == (3, 4) and labels.sum() == 16 and scores.tolist() == [0.0]
        np.testing.assert_allclose(poses[:, :3], np.array(g['R_gcransac']), atol=1e-5)
        np.testing.assert_allclose(poses[:, 3], np.array(g['t_gcransac']), atol=1e-3)


def test_pose6dscene_fixture(golden):
    """95 correspondences + GT pose (example_pnp.ipynb:77-79: 1.9e-5 deg / 1e-4 mm with the stale non-EPOS path;
    tolerance here 1e-3 deg / 1e-2 mm)."""
    g = golden('pose6dscene.json')
    c, K, gt = np.array(g['corrs']), np.array(g['K']), np.array(g['gt_pose'])
    for seed in range(3):
        poses, labels, _, st = pf.find6DPoses(c[:, :2], c[:, 2:], K, threshold=4.0, min_triangle_area=0.0, seed=seed,
                                              return_stats=True)
        assert poses.shape == (3, 4)
        assert rot_err_deg(poses[:, :3], gt[:, :3]) < 1e-3
        assert np.linalg.norm(poses[:, 3] - gt[:, 3]) < 1e-2


def test_tless_fixture(golden):
    """1 886 EPOS-like correspondences, 2 GT poses; the notebook records 1 model, 1.98 deg / 1.26 cm to GT #2."""
    g = golden('tless.json')
    c, K, gts = np.array(g['corrs']), np.array(g['K']), np.array(g['gt_poses'])
    # The reference itself is randomised (unseeded mt19937) and the object is near-symmetric: some streams lock on
    # GT #1 or on a flipped pose.  Required: most seeds reproduce the recorded answer (about 1.9-2.0 deg, 1.2-1.4 cm
    # to GT #2, ~390 inliers of 1 886).
    hits = []
    for seed in range(5):
        poses, labels, _, st = pf.find6DPoses(c[:, :2], c[:, 2:], K, threshold=4.0, min_triangle_area=0.0, seed=seed,
                                              return_stats=True)
        assert poses.shape == (3, 4) and st['iterations'] >= 400
        e = (rot_err_deg(poses[:, :3], gts[1][:, :3]), np.linalg.norm(poses[:, 3] - gts[1][:, 3]))
        hits.append(e[0] < 3.5 and e[1] < 20.0 and labels.sum() > 370)
    assert sum(hits) >= 4, hits


def _cut_energy(g, lam, labels):
    """Energy of a labeling under the terms handed to add_term1 / add_term2 (GCRANSAC.h:843-907)."""
    e = float(np.where(labels == 1, g['u1'], g['u0']).sum())
    a, b = labels[g['ex']], labels[g['ey']]
    return e + float(np.where((a == 0) & (b == 0), g['e00'], np.where(a != b, lam, 0.0)).sum())


def _assert_same_cut(lab, ref, g, lam):
    """Identical labelings, except on exact-arithmetic TIES: when the terminal capacity of a group of nodes equals the
    total capacity of its arcs, both labels are optimal and the reference's BK answer depends on the rounding of its own
    subtraction sequence.  The oracle (and the CUDA kernel) resolve ties as exact arithmetic does (saturated ->
    SOURCE); a mismatch is accepted only if both labelings are minimum cuts of the same energy."""
    if np.array_equal(lab, ref):
        return 0
    ea, eb = _cut_energy(g, lam, lab), _cut_energy(g, lam, ref)
    assert abs(ea - eb) < 1e-9 * max(1.0, abs(eb)), (ea, eb)
    assert (lab != ref).sum() <= 8 and (lab <= ref).all()       # ties resolve towards SOURCE (outlier)
    return int((lab != ref).sum())


def test_labeling_matches_reference_bk_maxflow(golden):
    """oracle labeling (Dinic + reverse BFS) == the reference's own BK max-flow + what_segment on the same energy."""
    if pf.ref_lib() is None:
        pytest.skip('oracle/_ref not built (reference tree absent)')
    ties = 0
    g = golden('tless.json')
    c, K, gts = np.array(g['corrs']), np.array(g['K']), np.array(g['gt_poses'])
    rng = np.random.default_rng(3)
    nbr = pf.neighbors(c[:, :2], c[:, 2:], K)
    checked = 0
    for trial in range(12):
        model = gts[trial % 2].copy()
        rv = rng.normal(size=3) * 0.01 * (trial // 2)
        dR, _ = pf.rodrigues_to_matrix(rv)
        model[:, :3] = dR @ model[:, :3]
        model[:, 3] += rng.normal(size=3) * 2.0 * (trial // 2)
        for lam in (0.1, 0.3):
            p = pf.default_params(spatial_coherence_weight=lam)
            lab = pf.labeling(c[:, :2], c[:, 2:], K, model, nbr=nbr, params=p)
            gr = pf.cut_graph(c[:, :2], c[:, 2:], K, model, nbr, params=p)
            ties += _assert_same_cut(lab, pf.ref_bk_labeling(gr, lam), gr, lam)
            checked += 1
            thr = pf.score(c[:, :2], c[:, 2:], K, model)['mask']
            if trial == 0 and lam == 0.1:
                assert 0 < lab.sum() < len(lab)
                assert (lab != thr).sum() < 0.2 * len(lab)      # the cut is a perturbation of plain thresholding
    assert checked == 24
    # random graphs as well.  lambda = 0.1 has a STRUCTURAL tie when an outlier node (terminal weight (1-lambda)*1
    # = 0.9) has exactly 9 incident edges of weight 0.1: both labels then have the same energy and BK's answer
    # depends on rounding (DESIGN.md "graph cut").  So 0.1 is exercised on graphs of degree < 9 (as produced by the
    # <= 5-neighbour graphs of this path) and denser graphs use tie-free lambdas.
    for trial in range(12):
        n = 300
        X, uv, K2, R, t = _random_scene(rng, n, noise=2.0)
        out = rng.random(n) < 0.4
        uv[out] += rng.uniform(-30, 30, (int(out.sum()), 2))
        model = np.concatenate([R, t[:, None]], 1)
        dense = trial % 2 == 0
        nb = [list(rng.choice(n, size=rng.integers(0, 9 if dense else 3), replace=False)) for _ in range(n)]
        for lam in ((0.23, 0.3) if dense else (0.1, 0.3)):
            p = pf.default_params(spatial_coherence_weight=lam)
            lab = pf.labeling(uv, X, K2, model, nbr=nb, params=p)
            gr = pf.cut_graph(uv, X, K2, model, nb, params=p)
            ties += _assert_same_cut(lab, pf.ref_bk_labeling(gr, lam), gr, lam)
    assert ties <= 12                                            # a handful of tie nodes over 48 graphs
    # the planted scene on which the tie first showed (an outlier with 9 incident edges at lambda = 0.1): here the
    # reference's BK lands on the exact-arithmetic answer as well
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'case_b2_o2.npz'))
    p = pf.default_params()
    nbr = pf.neighbors(d['c2'], d['c3'], d['K'], p)
    lab = pf.labeling(d['c2'], d['c3'], d['K'], d['model'], nbr=nbr, params=p)
    gr = pf.cut_graph(d['c2'], d['c3'], d['K'], d['model'], nbr, params=p)
    assert np.array_equal(lab, pf.ref_bk_labeling(gr, 0.1)) and lab.sum() == 1336


def test_neighbors_are_nearest_within_radius(golden):
    g = golden('tless.json')
    c, K = np.array(g['corrs'])[:400], np.array(g['K'])
    nbr = pf.neighbors(c[:, :2], c[:, 2:], K)
    q = np.concatenate([c[:, :2], 0.1 * c[:, 2:]], 1).astype(np.float32)
    d2 = ((q[:, None, :] - q[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    for i in range(len(c)):
        order = np.lexsort((np.arange(len(c)), d2[i]))
        exp = [j for j in order[:5] if d2[i, j] <= 400.0 * (1 + 1e-6)]
        got = [j for j in nbr[i] if j >= 0]
        assert len(got) == len(exp)
        assert np.allclose(sorted(d2[i, got]), sorted(d2[i, exp]), rtol=1e-5)


def test_grid_binned_neighbour_search_equals_all_pairs(golden):
    """The oracle's neighbour graph is built with a uniform (u, v) grid (sub-quadratic, like the reference's KD-tree);
    the lists must be identical to the all-pairs scan on dense, sparse and degenerate point sets."""
    g = golden('tless.json')
    c, K = np.array(g['corrs']), np.array(g['K'])
    rng = np.random.default_rng(0)
    cases = [(c[:, :2], c[:, 2:])]
    cases.append((4.0 * (rng.integers(0, 160, (3000, 2)) + 0.5), rng.uniform(-100, 100, (3000, 3))))     # dense pixel grid
    cases.append((rng.uniform(0, 640, (50, 2)), rng.uniform(-5, 5, (50, 3))))                             # sparse
    cases.append((np.tile([[100.0, 100.0]], (40, 1)), np.tile([[1.0, 2.0, 3.0]], (40, 1))))               # all identical
    cases.append((np.stack([np.linspace(0, 1e5, 300), np.zeros(300)], 1), rng.uniform(-5, 5, (300, 3))))  # huge extent
    for uv, X in cases:
        for k in (5, 8):
            p = pf.default_params(max_neighbors=k)
            assert np.array_equal(pf.neighbors(uv, X, K, p), pf.neighbors_bruteforce(uv, X, K, p))


def test_find6dposes_argument_checks():
    K = np.eye(3)
    with pytest.raises(ValueError):
        pf.find6DPoses(np.zeros((5, 3)), np.zeros((5, 3)), K)
    with pytest.raises(ValueError):
        pf.find6DPoses(np.zeros((2, 2)), np.zeros((2, 3)), K)
    with pytest.raises(ValueError):
        pf.find6DPoses(np.zeros((5, 2)), np.zeros((5, 3)), np.eye(4))


def test_no_consensus_returns_no_pose():
    rng = np.random.default_rng(5)
    x2d = rng.uniform(0, 640, (60, 2))
    x3d = rng.uniform(-50, 50, (60, 3))
    K = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]])
    poses, labels, scores, st = pf.find6DPoses(x2d, x3d, K, threshold=0.05, min_triangle_area=0.0, seed=0,
                                               return_stats=True)
    assert poses.shape[0] in (0, 3)
    if poses.shape[0] == 0:
        assert labels.sum() == 0 and st['found'] == 0
