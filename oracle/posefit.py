"""ORACLE (test infrastructure, never imported by the product): ctypes front-end of oracle/posefit.cpp, the CPU
restatement of pyprogressivex.find6DPoses (single-instance branch: GC-RANSAC + final LM).  See the header of
posefit.cpp for the reference file:line map and the list of restated third-party algorithms."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


class Params(C.Structure):
    _fields_ = [('threshold', C.c_double), ('spatial_coherence_weight', C.c_double),
                ('neighborhood_ball_radius', C.c_double), ('scaling_from_millimeters', C.c_double),
                ('min_triangle_area', C.c_double), ('min_coverage', C.c_double), ('confidence', C.c_double),
                ('max_iters', C.c_int), ('min_iters', C.c_int), ('min_iters_before_lo', C.c_int),
                ('max_lo_trials', C.c_int), ('max_graph_cuts', C.c_int), ('max_lsq_iters', C.c_int),
                ('max_unsuccessful', C.c_int), ('max_neighbors', C.c_int), ('apply_numerical_optimization', C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'libposefit_oracle.so')
        src = os.path.join(_HERE, 'posefit.cpp')
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(['make', '-s', '-C', _HERE, 'libposefit_oracle.so'])
        _LIB = C.CDLL(so)
        _LIB.ora_rng_u64.restype = C.c_ulonglong
        _LIB.ora_rng_u64.argtypes = [C.c_ulonglong] * 5
    return _LIB


def ref_lib():
    """The reference's own BK max-flow (oracle/_ref/libref_maxflow.so), or None when not built."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, '_ref', 'libref_maxflow.so')
        if not os.path.exists(so):
            return None
        _REF = C.CDLL(so)
    return _REF


_REF_GCO = None


def ref_gco_lib():
    """The reference's own alpha-expansion (oracle/_ref/libref_gco.so), or None when not built."""
    global _REF_GCO
    if _REF_GCO is None:
        so = os.path.join(_HERE, '_ref', 'libref_gco.so')
        if not os.path.exists(so):
            return None
        _REF_GCO = C.CDLL(so)
        _REF_GCO.ref_gco_expansion.restype = C.c_double
    return _REF_GCO


def ref_alpha_expansion(D, nbr, lam, label_cost, labels=None, max_iterations=1000):
    """alpha_expansion() through the reference's GCoptimization (None when oracle/_ref is not built)."""
    r = ref_gco_lib()
    if r is None:
        return None
    D = _d(D)
    n, L = D.shape
    off, idx = _csr(nbr)
    lab = np.zeros(n, np.int32) if labels is None else np.ascontiguousarray(labels, np.int32).copy()
    e = r.ref_gco_expansion(n, L, _p(D), _p(off, C.c_int), _p(idx, C.c_int), C.c_double(lam), C.c_double(label_cost),
                            _p(lab, C.c_int), int(labels is not None), int(max_iterations))
    return lab, float(e)


def default_params(**kw):
    p = Params()
    lib().ora_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def rng_u64(seed, stream, a, b, c):
    return int(lib().ora_rng_u64(seed, stream, a, b, c))


def unique_set(seed, stream, a, b, n, k):
    out = np.zeros(k, np.int32)
    ok = lib().ora_unique_set(C.c_ulonglong(seed), C.c_ulonglong(stream), C.c_ulonglong(a), C.c_ulonglong(b), n, k,
                              _p(out, C.c_int))
    return out if ok else None


def quartic(coeffs):
    """coeffs c0..c4 (ascending powers) -> real roots ascending."""
    c = _d(coeffs)
    r = np.zeros(4)
    n = lib().ora_quartic(_p(c), _p(r))
    return r[:n]


def points7(x2d, x3d, K):
    x2d, x3d, K = _d(x2d), _d(x3d), _d(K)
    Kinv = np.linalg.inv(K)
    n = x2d.shape[0]
    pts = np.zeros((n, 7))
    h = np.concatenate([x2d, np.ones((n, 1))], 1)
    pts[:, 0] = h @ Kinv[0]
    pts[:, 1] = h @ Kinv[1]
    pts[:, 2:5] = x3d
    pts[:, 5:7] = x2d
    return pts


def p3p(pts7, idx):
    pts7 = _d(pts7)
    idx = np.ascontiguousarray(idx, np.int32)
    m = np.zeros((4, 12))
    n = lib().ora_p3p(_p(pts7), _p(idx, C.c_int), _p(m))
    return m[:n].reshape(n, 3, 4)


def rodrigues_to_matrix(r):
    r = _d(r)
    R, J = np.zeros(9), np.zeros(27)
    lib().ora_rodrigues_to_matrix(_p(r), _p(R), _p(J))
    return R.reshape(3, 3), J.reshape(3, 9)


def matrix_to_rodrigues(R):
    R = _d(R)
    r = np.zeros(3)
    lib().ora_matrix_to_rodrigues(_p(R), _p(r))
    return r


def solvepnp(X, uv, guess=None):
    """cv2.solvePnP(X, uv, I, None, flags=SOLVEPNP_ITERATIVE[, guess]) -> (ok, rvec, tvec)."""
    X, uv = _d(X), _d(uv)
    param = np.zeros(6)
    if guess is not None:
        param[:3], param[3:] = guess[0], guess[1]
    ok = lib().ora_solvepnp(X.shape[0], _p(X), _p(uv), int(guess is not None), _p(param))
    return bool(ok), param[:3].copy(), param[3:].copy()


def neighbors(x2d, x3d, K, params=None):
    params = params or default_params()
    x2d, x3d, K = _d(x2d), _d(x3d), _d(K)
    n = x2d.shape[0]
    out = np.zeros((n, params.max_neighbors), np.int32)
    lib().ora_neighbors(n, _p(x2d), _p(x3d), _p(K), C.byref(params), _p(out, C.c_int))
    return out


def neighbors_bruteforce(x2d, x3d, K, params=None):
    """All-pairs form of neighbors() (same lists; kept to test the grid-binned search)."""
    params = params or default_params()
    x2d, x3d, K = _d(x2d), _d(x3d), _d(K)
    n = x2d.shape[0]
    out = np.zeros((n, params.max_neighbors), np.int32)
    lib().ora_neighbors_bruteforce(n, _p(x2d), _p(x3d), _p(K), C.byref(params), _p(out, C.c_int))
    return out


def _csr(nbr):
    """[-1 padded N x k] or list of lists -> (offsets, index)."""
    lists = [[int(j) for j in row if j >= 0] for row in nbr]
    off = np.zeros(len(lists) + 1, np.int32)
    off[1:] = np.cumsum([len(l) for l in lists])
    idx = np.array([j for l in lists for j in l] or [0], np.int32)
    return off, idx


def score(x2d, x3d, K, model, params=None, best_inl=0):
    params = params or default_params()
    x2d, x3d, K, model = _d(x2d), _d(x3d), _d(K), _d(model)
    n = x2d.shape[0]
    out = np.zeros(3, np.int64)
    mask = np.zeros(n, np.int32)
    lib().ora_score(n, _p(x2d), _p(x3d), _p(K), C.byref(params), _p(model), C.c_longlong(best_inl),
                    _p(out, C.c_longlong), _p(mask, C.c_int))
    return {'value': int(out[0]), 'inliers': int(out[1]), 'used_pixels': int(out[2]), 'mask': mask}


def generate_models(x2d, x3d, K, seed, pass_index, params=None):
    params = params or default_params()
    x2d, x3d, K = _d(x2d), _d(x3d), _d(K)
    m = np.zeros((4, 12))
    fails = C.c_int(0)
    sample = np.zeros(3, np.int32)
    n = lib().ora_generate_models(x2d.shape[0], _p(x2d), _p(x3d), _p(K), C.byref(params), C.c_ulonglong(seed),
                                  pass_index, _p(m), C.byref(fails), _p(sample, C.c_int))
    return m[:n].reshape(n, 3, 4), fails.value, sample


def labeling(x2d, x3d, K, model, nbr=None, params=None):
    params = params or default_params()
    x2d, x3d, K, model = _d(x2d), _d(x3d), _d(K), _d(model)
    n = x2d.shape[0]
    labels = np.zeros(n, np.int32)
    if nbr is None:
        lib().ora_labeling(n, _p(x2d), _p(x3d), _p(K), C.byref(params), _p(model), None, None, _p(labels, C.c_int))
    else:
        off, idx = _csr(nbr)
        lib().ora_labeling(n, _p(x2d), _p(x3d), _p(K), C.byref(params), _p(model), _p(off, C.c_int), _p(idx, C.c_int),
                           _p(labels, C.c_int))
    return labels


def cut_graph(x2d, x3d, K, model, nbr, params=None):
    params = params or default_params()
    x2d, x3d, K, model = _d(x2d), _d(x3d), _d(K), _d(model)
    n = x2d.shape[0]
    off, idx = _csr(nbr)
    cap = max(1, idx.size)
    tr, u0, u1 = np.zeros(n), np.zeros(n), np.zeros(n)
    ex, ey = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    cxy, cyx, e00 = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    E = lib().ora_cut_graph(n, _p(x2d), _p(x3d), _p(K), C.byref(params), _p(model), _p(off, C.c_int), _p(idx, C.c_int),
                            _p(tr), _p(ex, C.c_int), _p(ey, C.c_int), _p(cxy), _p(cyx), _p(u0), _p(u1), _p(e00), cap)
    assert E >= 0
    return {'tr': tr, 'u0': u0, 'u1': u1, 'ex': ex[:E].copy(), 'ey': ey[:E].copy(), 'cxy': cxy[:E].copy(),
            'cyx': cyx[:E].copy(), 'e00': e00[:E].copy()}


def ref_bk_labeling(graph, lam):
    """Labels from the reference's BK max-flow on the raw energy terms (None when oracle/_ref is not built)."""
    r = ref_lib()
    if r is None:
        return None
    n, E = graph['u0'].size, graph['ex'].size
    labels = np.zeros(n, np.int32)
    e01 = np.full(max(E, 1), lam)
    e11 = np.zeros(max(E, 1))
    en = C.c_double(0)
    r.ref_bk_labeling(n, _p(graph['u0']), _p(graph['u1']), E, _p(graph['ex'], C.c_int), _p(graph['ey'], C.c_int),
                      _p(_d(graph['e00'] if E else np.zeros(1))), _p(e01), _p(e01), _p(e11), _p(labels, C.c_int),
                      C.byref(en))
    return labels


def alpha_expansion(D, nbr, lam, label_cost, labels=None, max_iterations=1000):
    """GCoptimization-style alpha-expansion (standard cycles) on data costs D [n, L], neighbour LISTINGS nbr (list of
    lists; a mutual pair listed from both sides weighs twice), Potts weight lam and a uniform label cost.
    Returns (labels, energy)."""
    D = _d(D)
    n, L = D.shape
    off, idx = _csr(nbr)
    lab = np.zeros(n, np.int32) if labels is None else np.ascontiguousarray(labels, np.int32).copy()
    f = lib().ora_alpha_expansion
    f.restype = C.c_double
    e = f(n, L, _p(D), _p(off, C.c_int), _p(idx, C.c_int), C.c_double(lam), C.c_double(label_cost), _p(lab, C.c_int),
          int(max_iterations))
    return lab, float(e)


def find6DPoses(x1y1, x2y2z2, K, threshold=4.0, max_model_number=1, conf=0.5, proposal_engine_conf=1.0,
                spatial_coherence_weight=0.1, neighborhood_ball_radius=20.0, max_tanimoto_similarity=0.9,
                scaling_from_millimeters=0.1, min_triangle_area=100.0, min_coverage=0.5, max_iters=400,
                min_point_number=6, use_prosac=False, max_model_number_for_optimization=3,
                apply_numerical_optimization=True, log=False, seed=0, nbr=None, return_stats=False, max_neighbors=5):
    """Same signature as pyprogressivex.find6DPoses (bindings.cpp:9-28,133-152) + seed / nbr.
    max_model_number == 1: GC-RANSAC + final LM; otherwise Progressive-X (PEARL for 2..max_model_number_for_optimization
    instances, sequential propose-and-remove beyond that or for -1)."""
    x1y1, x2y2z2, K = _d(x1y1), _d(x2y2z2), _d(K)
    if x1y1.ndim != 2 or x1y1.shape[1] != 2 or x2y2z2.ndim != 2 or x2y2z2.shape[1] != 3:
        raise ValueError('x1y1 should be an array with dims [n,2], x2y2z2 [n,3]')
    n = x1y1.shape[0]
    if n < 3 or x2y2z2.shape[0] != n:
        raise ValueError('x1y1 and x2y2z2 should be the same size, n>=3')
    if K.shape != (3, 3):
        raise ValueError('K should be an array with dims [3,3]')
    p = default_params(threshold=threshold, spatial_coherence_weight=spatial_coherence_weight,
                       neighborhood_ball_radius=neighborhood_ball_radius,
                       scaling_from_millimeters=scaling_from_millimeters, min_triangle_area=min_triangle_area,
                       min_coverage=min_coverage, confidence=proposal_engine_conf, max_iters=max_iters,
                       max_neighbors=max_neighbors, apply_numerical_optimization=int(apply_numerical_optimization))
    if max_model_number != 1:
        # Progressive-X (progressivex_python.cpp:136-221).  proposal_engine_conf is not forwarded by the reference in this
        # branch (the proposal engine runs at MultiModelSettings::proposal_engine_confidence = 1.0, progressive_x.h:66,695).
        if max_model_number == 0 or max_model_number < -1:
            raise ValueError('max_model_number should be -1 or positive')
        cap = 64 if max_model_number < 0 else max(max_model_number, 1) * 2 + 2
        poses = np.zeros((cap, 12))
        labels = np.zeros(n, np.int32)
        scores = np.zeros(cap)
        stats = np.zeros(6, np.int32)
        if nbr is None:
            off = idx = None
        else:
            off, idx = _csr(nbr)
        m = lib().ora_find6dposes_multi(n, _p(x1y1), _p(x2y2z2), _p(K), C.byref(p), C.c_ulonglong(seed),
                                        int(max_model_number), int(max_model_number_for_optimization), C.c_double(conf),
                                        C.c_double(max_tanimoto_similarity), int(min_point_number),
                                        None if off is None else _p(off, C.c_int), None if idx is None else _p(idx, C.c_int),
                                        cap, _p(poses), _p(labels, C.c_int), _p(scores), _p(stats, C.c_int))
        assert m <= cap
        out = (poses[:m].reshape(3 * m, 4).copy(), labels, scores[:m].copy())
        if return_stats:
            out = out + (dict(zip(('proposals', 'accepted', 'ransac_iterations', 'pearl_iterations', 'expansion_moves',
                                   'sped_up'), stats.tolist())),)
        return out
    pose = np.zeros(12)
    labels = np.zeros(n, np.int32)
    stats = np.zeros(5, np.int32)
    if nbr is None:
        r = lib().ora_find6dposes(n, _p(x1y1), _p(x2y2z2), _p(K), C.byref(p), C.c_ulonglong(seed), None, None,
                                  _p(pose), _p(labels, C.c_int), _p(stats, C.c_int))
    else:
        off, idx = _csr(nbr)
        r = lib().ora_find6dposes(n, _p(x1y1), _p(x2y2z2), _p(K), C.byref(p), C.c_ulonglong(seed), _p(off, C.c_int),
                                  _p(idx, C.c_int), _p(pose), _p(labels, C.c_int), _p(stats, C.c_int))
    poses = pose.reshape(3, 4) if r else np.zeros((0, 4))
    scores = np.zeros(1 if r else 0)            # RANSACStatistics::score is never written (statistics.h:67)
    out = (poses, labels, scores)
    if return_stats:
        out = out + (dict(zip(('iterations', 'graph_cuts', 'lo_runs', 'passes', 'found'), stats.tolist())),)
    return out
