"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/epos_b200.h
declares (no compute calls), the Python shims validate arguments like the reference binding, and the product never
imports the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'epos_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(epos_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from epos_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(_lib.exported_symbols()) == set(names)
    l = _lib.lib()
    assert l.epos_version() >= 1 and l.epos_compiled_arch() == 100 and l.epos_fit_max_points() == 4096
    assert l.epos_fit_workspace_bytes(4, 4096, None) > 0 and l.epos_corresp_workspace_bytes(2, 3, 120, 160) > 0


def test_sass_has_blackwell_instructions():
    """tcgen05.mma / tcgen05.ld / TMA appear as UTC*MMA / LDTM / UTMALDG in the built library."""
    import shutil
    import subprocess
    from epos_b200 import _lib
    if not shutil.which('cuobjdump'):
        pytest.skip('cuobjdump not on PATH')
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'UTCHMMA' in sass or 'UTCMMA' in sass
    assert 'LDTM' in sass and 'UTMALDG' in sass


def test_find6dposes_shim_argument_errors():
    from epos_b200 import posefit
    K = np.eye(3)
    with pytest.raises(ValueError):
        posefit.find6DPoses(np.zeros((5, 3)), np.zeros((5, 3)), K, max_model_number=1)
    with pytest.raises(ValueError):
        posefit.find6DPoses(np.zeros((2, 2)), np.zeros((2, 3)), K, max_model_number=1)
    with pytest.raises(ValueError):
        posefit.find6DPoses(np.zeros((5, 2)), np.zeros((4, 3)), K, max_model_number=1)
    with pytest.raises(ValueError):
        posefit.find6DPoses(np.zeros((5, 2)), np.zeros((5, 3)), np.eye(4), max_model_number=1)
    with pytest.raises(ValueError):
        posefit.find6DPoses(np.zeros((5, 2)), np.zeros((5, 3)), K, max_model_number=0)
    with pytest.raises(ValueError):                                          # PEARL holds at most 5 instances + the proposal
        posefit.find6DPoses(np.zeros((5, 2)), np.zeros((5, 3)), K, max_model_number=2, max_model_number_for_optimization=9)
    with pytest.raises(NotImplementedError):
        posefit.find6DPoses(np.zeros((5, 2)), np.zeros((5, 3)), K, max_model_number=1, proposal_engine_conf=0.9)


def test_fit_params_defaults_match_reference_flags():
    """scripts/infer.py:76-120 defaults and progressivex_python.cpp:223-234 / settings.h:68-88."""
    from epos_b200 import posefit
    p = posefit.default_params()
    assert (p.threshold, p.spatial_coherence_weight, p.neighborhood_ball_radius) == (4.0, 0.1, 20.0)
    assert (p.scaling_from_millimeters, p.min_triangle_area, p.min_coverage) == (0.1, 0.0, 0.5)
    assert (p.max_iters, p.min_iters, p.min_iters_before_lo, p.max_lo_trials) == (400, 10, 20, 20)
    assert (p.max_graph_cuts, p.max_lsq_iters, p.max_unsuccessful, p.max_neighbors) == (10, 10, 100, 5)
    from oracle import posefit as opf
    q = opf.default_params()
    for k in ('threshold', 'spatial_coherence_weight', 'neighborhood_ball_radius', 'scaling_from_millimeters',
              'min_triangle_area', 'min_coverage', 'max_iters', 'min_iters', 'min_iters_before_lo', 'max_lo_trials',
              'max_graph_cuts', 'max_lsq_iters', 'max_unsuccessful', 'max_neighbors', 'apply_numerical_optimization'):
        assert getattr(p, k) == getattr(q, k), k


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'epos_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')) and f != 'smoke.py':
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M), f


def test_weight_blob_roundtrip_and_sharding():
    from epos_b200 import dist as d, weights as W
    w = W.random_init(2, 4, seed=1, bn='random')
    blob = d.pack_weights(w, 2, 4)
    assert blob.size == d.blob_size(2, 4)
    u = d.unpack_weights(blob, 2, 4)
    assert set(u) == set(w) and all(np.array_equal(u[k], w[k]) for k in w)
    assert [d.shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [d.shard_range(8, 8, r) for r in range(8)] == [(r, r + 1) for r in range(8)]
