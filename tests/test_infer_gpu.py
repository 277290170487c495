"""scripts/infer.py end to end on the GPU: the synthetic CNN path, the LOCALIZATION path with two planted instances
per object (Progressive-X + PEARL through Engine / BatchFitter), and DETECTION (all instances), each writing a BOP CSV
that parses with the bop_toolkit reader."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _infer():
    spec = importlib.util.spec_from_file_location('epos_infer_cli', os.path.join(ROOT, 'scripts', 'infer.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_infer_cli_synthetic_images(tmp_path):
    from epos_b200 import bop_io
    m = _infer()
    res = m.main(['--num_images', '2', '--batch_size', '2', '--num_objs', '3', '--num_frags', '16', '--head_std', '30',
                  '--infer_dir', str(tmp_path), '--infer_name', 't', '--vis'])
    back = bop_io.load_bop_results(str(tmp_path / 'estimated-poses_t.csv'))
    assert len(back) == len(res)
    for a, b in zip(res, back):
        assert a['obj_id'] == b['obj_id'] and a['im_id'] == b['im_id'] and np.array_equal(a['R'], b['R'])
    assert sorted(os.listdir(str(tmp_path / 'vis'))) == ['000000_grid.jpg', '000001_grid.jpg']


def test_infer_cli_planted_two_instances_per_object(tmp_path):
    from epos_b200 import bop_io, synthetic
    m = _infer()
    args = ['--num_images', '2', '--batch_size', '2', '--num_objs', '3', '--num_frags', '64', '--planted',
            '--instances_per_object', '2', '--infer_dir', str(tmp_path), '--seed', '3']
    res = m.main(args)
    back = bop_io.load_bop_results(str(tmp_path / 'estimated-poses.csv'))
    assert len(back) == len(res) >= 8
    # the planted ground truth: every recovered pose of an object is close to one of its planted instances, and most
    # planted instances are recovered (two spheres of one object may overlap)
    K = synthetic.default_K()
    store = synthetic.model_store(3, 64)
    hit = tot = 0
    for b in range(2):
        _, _, _, gt = synthetic.planted_maps(2, 3, 64, store, K, seed=3, objs_per_image=3, instances_per_object=2)
        for oid, insts in gt[b].items():
            est = [r for r in res if r['im_id'] == b and r['obj_id'] == oid]
            assert 1 <= len(est) <= 2
            for R, t in insts:
                tot += 1
                hit += any(np.abs(e['R'] - R).max() < 3e-2 and np.linalg.norm(e['t'].ravel() - t) < 8.0 for e in est)
            assert all(r['score'] > 10 for r in est)                 # Progressive-X scores = support of the instance
    assert hit >= tot - 2, (hit, tot)


def test_infer_cli_detection_returns_all_instances(tmp_path):
    m = _infer()
    res = m.main(['--num_images', '1', '--batch_size', '1', '--num_objs', '2', '--num_frags', '64', '--planted',
                  '--instances_per_object', '2', '--task_type', 'detection', '--infer_dir', str(tmp_path), '--seed', '5',
                  '--save_estimates', 'false'])
    assert not os.path.exists(str(tmp_path / 'estimated-poses.csv'))
    per_obj = {}
    for r in res:
        per_obj[r['obj_id']] = per_obj.get(r['obj_id'], 0) + 1
    assert per_obj and all(2 <= n <= 32 for n in per_obj.values()), per_obj
    assert all(r['score'] == 0.0 for r in res)                        # spedUpFitting never writes a score
