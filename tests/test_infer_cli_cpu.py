"""CPU tests of the scripts/infer.py front end: flags that cannot be honoured stop the run (no silent fallbacks), the
num_instances rule of /root/reference/scripts/infer.py:462-468, and the BOP CSV written by save_bop_results parses
with the reader of bop_toolkit (inout.py:220-262) back to the same poses."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _infer():
    spec = importlib.util.spec_from_file_location('epos_infer_cli', os.path.join(ROOT, 'scripts', 'infer.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_flag_names_and_defaults():
    a = _infer().parse_args([])
    assert (a.task_type, a.fitting_method, a.inlier_thresh, a.neighbour_max_dist, a.min_hypothesis_quality) == \
        ('localization', 'progressive_x', 4.0, 20.0, 0.5)
    assert (a.required_progx_confidence, a.required_ransac_confidence, a.min_triangle_area, a.use_prosac) == (0.5, 1.0, 0.0, False)
    assert (a.max_model_number_for_pearl, a.spatial_coherence_weight, a.scaling_from_millimeters) == (5, 0.1, 0.1)
    assert (a.max_tanimoto_similarity, a.max_instances_to_fit, a.max_fitting_iterations, a.vis) == (0.9, None, 400, False)


@pytest.mark.parametrize('argv', [['--use_prosac'], ['--fitting_method', 'opencv_ransac'], ['--project_to_surface'],
                                  ['--required_ransac_confidence', '0.9'], ['--max_model_number_for_pearl', '9'],
                                  ['--instances_per_object', '0']])
def test_unsupported_flags_stop_the_run(argv):
    with pytest.raises(SystemExit):
        _infer().parse_args(argv)


def test_num_instances_rule():
    m = _infer()
    assert (m.num_instances_for(m.parse_args([]), 2, 3) == 1).all()
    assert (m.num_instances_for(m.parse_args(['--instances_per_object', '3']), 2, 3) == 3).all()
    assert (m.num_instances_for(m.parse_args(['--instances_per_object', '3', '--max_instances_to_fit', '2']), 1, 4) == 2).all()
    assert (m.num_instances_for(m.parse_args(['--task_type', 'detection']), 1, 4) == -1).all()
    # min(-1, k) = -1: the cap does not apply to DETECTION, as in the reference
    assert (m.num_instances_for(m.parse_args(['--task_type', 'detection', '--max_instances_to_fit', '2']), 1, 4) == -1).all()


def test_bop_csv_round_trip(tmp_path):
    from epos_b200 import bop_io
    rng = np.random.default_rng(0)
    res = []
    for i in range(5):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        res.append({'scene_id': 3, 'im_id': 10 + i, 'obj_id': 1 + i % 3, 'score': float(rng.uniform()), 'R': q,
                    't': rng.normal(size=(3, 1)) * 500, 'time': 0.0123 * (i + 1)})
    res.append({'scene_id': 0, 'im_id': 0, 'obj_id': 7, 'score': 0.0, 'R': np.eye(3), 't': np.zeros((3, 1))})   # no time -> -1
    path = str(tmp_path / 'estimated-poses.csv')
    bop_io.save_bop_results(path, res)
    lines = open(path).read().split('\n')
    assert lines[0] == 'scene_id,im_id,obj_id,score,R,t,time' and len(lines) == 7 and not lines[-1].endswith('\n')
    assert len(lines[1].split(',')) == 7 and len(lines[1].split(',')[4].split()) == 9 and len(lines[1].split(',')[5].split()) == 3
    back = bop_io.load_bop_results(path)
    assert len(back) == 6
    for a, b in zip(res, back):
        assert (a['scene_id'], a['im_id'], a['obj_id']) == (b['scene_id'], b['im_id'], b['obj_id'])
        assert a['score'] == b['score'] and np.array_equal(np.asarray(a['R']), b['R']) and np.array_equal(a['t'], b['t'])
        assert b['time'] == a.get('time', -1)
    with pytest.raises(ValueError):
        bop_io.save_bop_results(path, res, version='bop18')


def test_visualize_writes_a_grid(tmp_path):
    cv2 = pytest.importorskip('cv2')
    m = _infer()
    img = np.random.default_rng(0).integers(0, 255, (480, 640, 3)).astype(np.float32)
    labels = np.random.default_rng(1).integers(0, 5, (120, 160))
    K = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]])
    poses = [{'obj_id': 2, 'R': np.eye(3), 't': np.array([[0.0], [0.0], [800.0]])},
             {'obj_id': 3, 'R': np.eye(3), 't': np.array([[0.0], [0.0], [-5.0]])}]          # behind the camera: skipped
    path = str(tmp_path / 'grid.jpg')
    m.visualize(path, img, labels, poses, K)
    g = cv2.imread(path)
    assert g is not None and g.shape == (225, 900, 3)
