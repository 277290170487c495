"""Generates tests/golden/*.json from the reference tree (run in the build container only).

Everything extracted here is DATA the reference's own tests / example notebooks hold for the
inference hot path; no reference source code is copied.

  slim_conv2d_same.json   expected matrices of ResnetUtilsTest
                          (/root/reference/external/slim/nets/resnet_v1_test.py:58-149)
  pnp16.json              16 correspondences + K + recorded R,t
                          (.../graph-cut-ransac/examples/example_pnp.ipynb cells 4-7)
  pose6dscene.json        95 correspondences + K + GT pose  (.../examples/img/pose6dscene*)
  tless.json              1886 correspondences + K + 2 GT poses (external/progressive-x/examples/img/tless*)
  cv2_solvepnp.json       outputs of cv2.solvePnP(SOLVEPNP_ITERATIVE) (cv2 wheel in this image; the
                          reference calls OpenCV 3.4.2 which is not installed) on seeded point sets,
                          used to pin the oracle's restatement of the non-minimal solver.

Usage: python tests/golden/make_golden.py
"""
import json
import os
import re

import numpy as np

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))
PX = REF + '/external/progressive-x'
GC = PX + '/graph-cut-ransac'


def dump(name, obj):
    with open(os.path.join(OUT, name), 'w') as f:
        json.dump(obj, f)
    print('wrote', name)


def slim():
    src = open(REF + '/external/slim/nets/resnet_v1_test.py').read()

    def grab(fn, var):
        body = src[src.index('def ' + fn):]
        m = re.search(var + r' = tf\.cast\((\[\[.*?\]\]), tf\.float32\)', body, re.S)
        return json.loads(re.sub(r'\s+', '', m.group(1)))

    d = {}
    for fn, n in (('testConv2DSameEven', 4), ('testConv2DSameOdd', 5)):
        d[fn] = {'n': n, 'y1': grab(fn, 'y1_expected'), 'y2': grab(fn, 'y2_expected'),
                 'y4': grab(fn, 'y4_expected') if n == 4 else None}
    d['subsample3'] = [0, 2, 6, 8]
    d['subsample4'] = [0, 2, 8, 10]
    dump('slim_conv2d_same.json', d)


def pnp16():
    nb = json.load(open(GC + '/examples/example_pnp.ipynb'))
    cells = [''.join(c['source']) for c in nb['cells']]
    src = [c for c in cells if c.startswith('corrs = np.array')][0]
    rows = re.findall(r'\[([-0-9., e]+)\]', src)
    corrs = [[float(v) for v in r.split(',')] for r in rows]
    assert len(corrs) == 16
    out = [c for c in nb['cells'] if 'find6DPoseEPOS' in ''.join(c['source'])][0]['outputs'][0]['text']
    txt = ''.join(out)
    nums = [float(v) for v in re.findall(r'-?\d+\.\d+(?:e-?\d+)?', txt)]
    R = np.array(nums[:9]).reshape(3, 3).tolist()
    t = nums[9:12]
    dump('pnp16.json', {'corrs': corrs, 'K': [[1066.778, 0.0, 312.9869], [0.0, 1067.487, 241.3109], [0, 0, 1]],
                        'threshold_px': 4.0, 'R_gcransac': R, 't_gcransac': t,
                        'R_cv_ransac': [[0.7091456, 0.70483857, 0.01775135], [0.25790341, -0.23588271, -0.93693392],
                                        [-0.65619993, 0.6690007, -0.34905546]],
                        't_cv_ransac': [-86.4400753, 33.35438989, 777.05154377]})


def scenes():
    d = GC + '/examples/img/'
    dump('pose6dscene.json', {'corrs': np.loadtxt(d + 'pose6dscene_points.txt').tolist(),
                              'K': np.loadtxt(d + 'pose6dscene.K').tolist(),
                              'gt_pose': np.loadtxt(d + 'pose6dscene_gt.txt').reshape(3, 4).tolist()})
    d = PX + '/examples/img/'
    dump('tless.json', {'corrs': np.loadtxt(d + 'tless.txt').tolist(),
                        'K': np.loadtxt(d + 'tless_intrinsics.txt').tolist(),
                        'gt_poses': np.loadtxt(d + 'tless_poses.txt').reshape(-1, 3, 4).tolist(),
                        'recorded': {'models': 1, 'rot_err_deg_vs_gt2': 1.98, 'trans_err_cm_vs_gt2': 1.26,
                                     'seconds': 0.1157}})


def cv2_solvepnp():
    import cv2
    rng = np.random.Generator(np.random.PCG64(7))
    cases = []
    for n in (4, 6, 8, 21, 21, 60, 200):
        for planar in (False, True):
            if n < 6 and not planar:
                continue
            X = rng.uniform(-60, 60, size=(n, 3))
            if planar:
                X[:, 2] = 0.0
                Rp = cv2.Rodrigues(rng.normal(size=3))[0]
                X = X @ Rp.T + rng.uniform(-10, 10, 3)
            rv = rng.normal(size=3) * 0.8
            tv = np.array([rng.uniform(-100, 100), rng.uniform(-100, 100), rng.uniform(500, 1200)])
            R = cv2.Rodrigues(rv)[0]
            Xc = X @ R.T + tv
            uv = Xc[:, :2] / Xc[:, 2:3] + rng.normal(0, 1.5e-3, size=(n, 2))
            ok, r0, t0 = cv2.solvePnP(X, uv, np.eye(3), None, flags=cv2.SOLVEPNP_ITERATIVE)
            g_r = rv + rng.normal(0, 0.05, 3)
            g_t = tv + rng.normal(0, 5, 3)
            ok2, r1, t1 = cv2.solvePnP(X, uv, np.eye(3), None, g_r.reshape(3, 1).copy(), g_t.reshape(3, 1).copy(),
                                       True, cv2.SOLVEPNP_ITERATIVE)
            cases.append({'X': X.tolist(), 'uv': uv.tolist(), 'planar': planar, 'ok': bool(ok),
                          'rvec': r0.ravel().tolist(), 'tvec': t0.ravel().tolist(),
                          'R': cv2.Rodrigues(r0)[0].tolist(),
                          'guess_rvec': g_r.tolist(), 'guess_tvec': g_t.tolist(),
                          'rvec_guess': r1.ravel().tolist(), 'tvec_guess': t1.ravel().tolist()})
    # Rodrigues round trips
    rod = []
    for _ in range(20):
        rv = rng.normal(size=3) * rng.choice([1e-9, 1e-3, 0.5, 2.0, 3.1])
        R, J = cv2.Rodrigues(rv)
        rod.append({'rvec': rv.tolist(), 'R': R.tolist(), 'J': J.tolist(), 'back': cv2.Rodrigues(R)[0].ravel().tolist()})
    dump('cv2_solvepnp.json', {'cv2_version': cv2.__version__, 'cases': cases, 'rodrigues': rod})


def tie_case():
    """case_b2_o2.npz: problem (image 2, object 2) of the planted 8 x 21 batch of tests/test_pose_gpu.py, the first scene
    on which a graph-cut TIE showed (an outlier node with exactly (1-lambda)/lambda = 9 incident edges): correspondences,
    K, the model that is labelled (best P3P hypothesis of the first 11 passes) and the RANSAC stream key."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from epos_b200 import synthetic
    from oracle import pipeline, posefit as pf
    O, F, B = 21, 64, 8
    store, K = synthetic.model_store(O, F), synthetic.default_K()
    oc, fc, fl, _ = synthetic.planted_maps(B, O, F, store, K, seed=21, objs_per_image=5)
    pp = pipeline.PostProcess(O, F, seed=9, model_store=store, K=K, max_correspondences=2048)
    d = pp.corresp({'pred_obj_conf': oc, 'pred_frag_conf': fc, 'pred_frag_loc': fl}, 2)[2]
    seed = (9 << 32) + 2 * 21 + 1
    p = pf.default_params()
    best, bv, bi = None, 0, 0
    for ps in range(11):
        models, _, _ = pf.generate_models(d['coord_2d'], d['coord_3d'], K, seed, ps, p)
        for m in models:
            sc = pf.score(d['coord_2d'], d['coord_3d'], K, m, p, best_inl=bi)
            if bv < sc['value']:
                bv, bi, best = sc['value'], sc['inliers'], m.copy()
    np.savez_compressed(os.path.join(OUT, "case_b2_o2.npz"), c2=d['coord_2d'], c3=d['coord_3d'], K=K, model=best,
                        seed=np.int64(seed))


if __name__ == '__main__':
    tie_case()
    slim()
    pnp16()
    scenes()
    cv2_solvepnp()
