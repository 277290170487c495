#!/usr/bin/env python
"""Entry point with the flags of /root/reference/scripts/infer.py (:37-146) for the B200 path.

The reference reads TFRecords and a TF checkpoint; this build has neither TensorFlow nor the BOP datasets, so images
are synthetic (--synthetic, the default and only input mode) and weights are random-init or an .npz keyed by the TF
variable names (--weights, see scripts/convert_checkpoint.py).  Per-image flow, timing keys and the BOP CSV row format
follow process_image (infer.py:348-554), main (:712-760) and bop_toolkit inout.save_bop_results (inout.py:265-294).

Task types (infer.py:383-392,462-468): LOCALIZATION fits, per annotated object, as many instances as the annotation
holds (1: GC-RANSAC + final LM; 2..max_model_number_for_pearl: Progressive-X with PEARL; more: sequential
propose-and-remove).  Synthetic images carry no annotation, so --instances_per_object N stands in for the ground-truth
instance count of every object (default 1).  DETECTION asks for all instances (num_instances = -1); the reference's loop
for that case never terminates (progressive_x.h:280), here it is bounded (include/epos_b200.h, epos_fit_poses_multi).

Flags that cannot be honoured stop the run instead of being ignored: --use_prosac (the reference only re-orders the
correspondences for a sampler its C++ never selects, progressivex_python.cpp:55), --fitting_method opencv_ransac,
--project_to_surface (needs igl), --required_ransac_confidence != 1.

  python scripts/infer.py --num_images 16 --batch_size 8 --num_objs 21 --num_frags 64 [--world_size N via torchrun]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    # flags of the reference (same names and defaults)
    ap.add_argument('--model', default='synthetic')
    ap.add_argument('--task_type', default='localization', choices=['localization', 'detection'])
    ap.add_argument('--infer_name', default=None)
    ap.add_argument('--fitting_method', default='progressive_x', choices=['progressive_x', 'opencv_ransac'])
    ap.add_argument('--inlier_thresh', type=float, default=4.0)
    ap.add_argument('--neighbour_max_dist', type=float, default=20.0)
    ap.add_argument('--min_hypothesis_quality', type=float, default=0.5)
    ap.add_argument('--required_progx_confidence', type=float, default=0.5)
    ap.add_argument('--required_ransac_confidence', type=float, default=1.0)
    ap.add_argument('--min_triangle_area', type=float, default=0.0)
    ap.add_argument('--use_prosac', action='store_true')
    ap.add_argument('--project_to_surface', action='store_true')
    ap.add_argument('--max_model_number_for_pearl', type=int, default=5)
    ap.add_argument('--spatial_coherence_weight', type=float, default=0.1)
    ap.add_argument('--scaling_from_millimeters', type=float, default=0.1)
    ap.add_argument('--max_tanimoto_similarity', type=float, default=0.9)
    ap.add_argument('--max_correspondences', type=int, default=4096,
                    help='reference default None (unbounded); this build keeps at most 4096 per object')
    ap.add_argument('--max_instances_to_fit', type=int, default=None)
    ap.add_argument('--max_fitting_iterations', type=int, default=400)
    ap.add_argument('--corr_min_obj_conf', type=float, default=0.1)
    ap.add_argument('--corr_min_frag_rel_conf', type=float, default=0.5)
    ap.add_argument('--save_estimates', type=lambda s: s.lower() not in ('0', 'false', 'no'), default=True)
    ap.add_argument('--vis', action='store_true', help='per-image grid: input, predicted object labels, pose axes')
    ap.add_argument('--infer_dir', default=os.path.join(ROOT, 'gpurun_out', 'infer'))
    # additions of this build
    ap.add_argument('--synthetic', action='store_true', default=True)
    ap.add_argument('--weights', default=None, help='.npz of TF-named variables (epos_b200.weights.save_npz)')
    ap.add_argument('--num_images', type=int, default=8)
    ap.add_argument('--batch_size', type=int, default=8)
    ap.add_argument('--num_objs', type=int, default=21)
    ap.add_argument('--num_frags', type=int, default=64)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--head_std', type=float, default=300.0, help='logit initialiser stddev of the random-init heads')
    ap.add_argument('--model_variant', default='xception_65', choices=['xception_65', 'resnet_v1_50_beta'],
                    help='common.py:117-119 flag of the reference (backbone)')
    ap.add_argument('--instances_per_object', type=int, default=1,
                    help='LOCALIZATION: stands in for the annotated instance count of every object (synthetic input)')
    ap.add_argument('--planted', action='store_true',
                    help='replace the network outputs by planted maps with known poses (pose fitting only; synthetic.py)')
    args = ap.parse_args(argv)
    check_flags(args)
    return args


def check_flags(args):
    """Flag combinations this build cannot honour stop the run (no silent fallbacks)."""
    if args.required_ransac_confidence != 1.0:
        raise SystemExit('--required_ransac_confidence must be 1.0 (the iteration bound is max_fitting_iterations)')
    if args.use_prosac:
        raise SystemExit('--use_prosac is not supported: the reference only sorts the correspondences by confidence for a '
                         'PROSAC sampler that its C++ never instantiates (progressivex_python.cpp:55,120-126)')
    if args.fitting_method != 'progressive_x':
        raise SystemExit('--fitting_method opencv_ransac (cv2.solvePnPRansac, infer.py:505-528) is not built')
    if args.project_to_surface:
        raise SystemExit('--project_to_surface needs igl (infer.py:59-61) and is not built')
    if args.max_model_number_for_pearl > 5 or args.max_model_number_for_pearl < 1:
        raise SystemExit('--max_model_number_for_pearl must be in 1..5')
    if args.instances_per_object < 1:
        raise SystemExit('--instances_per_object must be >= 1')
    if args.max_instances_to_fit is not None and args.max_instances_to_fit < 1:
        raise SystemExit('--max_instances_to_fit must be >= 1')


def num_instances_for(args, B, J):
    """infer.py:462-468: GT instance count (LOCALIZATION) or -1 (DETECTION), capped by --max_instances_to_fit."""
    import numpy as np
    n = args.instances_per_object if args.task_type == 'localization' else -1
    if args.max_instances_to_fit is not None:
        n = min(n, args.max_instances_to_fit)
    return np.full((B, J), n, np.int32)


def colorize_label_map(labels):
    import numpy as np
    palette = np.array([[0, 0, 0]] + [[(37 * i) % 256, (91 * i + 60) % 256, (173 * i + 120) % 256] for i in range(1, 256)], np.uint8)
    return palette[np.asarray(labels) % 256]


def visualize(path, image, obj_labels, poses, K):
    """Grid of the reference's visualisation (infer.py:150-291) that needs no renderer: input image, predicted object
    labels, and the estimated poses drawn as projected object axes (the reference renders the object models with
    bop_renderer, which is out of scope)."""
    import cv2
    import numpy as np
    tile = (300, 225)
    rgb = np.clip(image, 0, 255).astype(np.uint8)
    lab = cv2.resize(colorize_label_map(obj_labels), (rgb.shape[1], rgb.shape[0]), interpolation=cv2.INTER_NEAREST)
    over = rgb.copy()
    for p in poses:
        R, t = p['R'], p['t'].reshape(3)
        pts = np.array([[0, 0, 0], [60, 0, 0], [0, 60, 0], [0, 0, 60]], np.float64) @ R.T + t
        if (pts[:, 2] <= 1e-6).any():
            continue
        uv = (pts[:, :2] / pts[:, 2:3]) * np.array([K[0, 0], K[1, 1]]) + np.array([K[0, 2], K[1, 2]])
        if not np.isfinite(uv).all() or np.abs(uv).max() > 1e5:
            continue
        o = tuple(int(v) for v in uv[0])
        for k, col in ((1, (255, 0, 0)), (2, (0, 255, 0)), (3, (0, 0, 255))):
            cv2.line(over, o, tuple(int(v) for v in uv[k]), col, 2)
        cv2.putText(over, str(p['obj_id']), o, cv2.FONT_HERSHEY_SIMPLEX, 0.5, (255, 255, 255), 1)
    tiles = [cv2.resize(x, tile) for x in (rgb, over, lab)]
    for im, name in zip(tiles, ('input', 'pred poses', 'predicted obj labels')):
        cv2.putText(im, name, (5, 15), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (204, 204, 204), 1)
    cv2.imwrite(path, cv2.cvtColor(np.concatenate(tiles, 1), cv2.COLOR_RGB2BGR))


def main(argv=None):
    args = parse_args(argv)
    import numpy as np
    import torch
    import torch.distributed as dist
    from epos_b200 import bop_io, dist as edist, engine, model, posefit, synthetic, weights as W
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    O, F = args.num_objs, args.num_frags
    w = None
    if rank == 0:
        w = W.load_npz(args.weights) if args.weights else W.random_init(O, F, seed=args.seed, logits_std=args.head_std,
                                                                        model_variant=args.model_variant)
    w = edist.broadcast_weights(w, O, F, dev, world, rank, model_variant=args.model_variant)
    store = synthetic.model_store(O, F)
    K = synthetic.default_K()
    params = posefit.default_params(threshold=args.inlier_thresh, neighborhood_ball_radius=args.neighbour_max_dist,
                                    min_coverage=args.min_hypothesis_quality, min_triangle_area=args.min_triangle_area,
                                    spatial_coherence_weight=args.spatial_coherence_weight,
                                    scaling_from_millimeters=args.scaling_from_millimeters,
                                    max_iters=args.max_fitting_iterations)
    mparams = posefit.multi_params(args.max_model_number_for_pearl, 6, args.required_progx_confidence,
                                   args.max_tanimoto_similarity)
    eng = engine.Engine(w, O, F, dev, stages=engine.STAGES_FULL, model_store=store, K=K, fit_params=params,
                        max_correspondences=args.max_correspondences, seed=args.seed,
                        min_obj_conf=args.corr_min_obj_conf, min_frag_rel_conf=args.corr_min_frag_rel_conf,
                        model_options=model.ModelOptions(W.head_channels(O, F), model_variant=args.model_variant),
                        multi_params=mparams)
    lo, hi = edist.shard_range(args.num_images, world, rank)
    results = []
    ids = store.dp_model['obj_ids']
    J = len(ids)
    vis_dir = os.path.join(args.infer_dir, 'vis')
    if args.vis and rank == 0:
        os.makedirs(vis_dir, exist_ok=True)
    for b0 in range(lo, hi, args.batch_size):
        n = min(args.batch_size, hi - b0)
        imgs_np = np.concatenate([W.synthetic_images(1, seed=10000 + i) for i in range(b0, b0 + n)])
        imgs = torch.from_numpy(imgs_np).pin_memory()
        ninst = num_instances_for(args, n, J)
        torch.cuda.synchronize()
        t0 = time.time()
        if args.planted:
            # known poses: planted network outputs (instances_per_object instances of every visible object)
            oc, fc, fl, gt = synthetic.planted_maps(n, O, F, store, K, seed=args.seed + b0, objs_per_image=min(3, O),
                                                    instances_per_object=args.instances_per_object)
            recs_t = eng._fitter.fit_maps(torch.from_numpy(oc).to(dev), torch.from_numpy(fc).to(dev),
                                          torch.from_numpy(fl).to(dev), num_instances=ninst)
            out = {'poses': recs_t, 'multi': eng._fitter.multi, model.PRED_OBJ_LABEL: torch.from_numpy(oc.argmax(-1)).to(dev)}
        else:
            out = eng.run_device(imgs.to(dev, non_blocking=True), num_instances=ninst)
        torch.cuda.synchronize()
        recs = out['poses'].cpu().numpy()                     # [n, J, 16]
        multi = out.get('multi')
        total = time.time() - t0
        print('Images: {}-{}, total time: {:.3f} s ({:.1f} images/s)'.format(b0, b0 + n - 1, total, n / total), flush=True)
        per_image = [[] for _ in range(n)]
        multi_index = {}
        if multi is not None:
            mp_, ms_, mc_ = (multi[k].cpu().numpy() for k in ('poses', 'scores', 'counts'))
            multi_index = {bj: q for q, bj in enumerate(multi['index'])}
        for i in range(n):
            for j, oid in enumerate(ids):
                if (i, j) in multi_index:                     # Progressive-X: one pose per instance (infer.py:490-503)
                    q = multi_index[(i, j)]
                    for k in range(max(int(mc_[q]), 0)):
                        P = mp_[q, k].reshape(3, 4)
                        per_image[i].append({'scene_id': 0, 'im_id': b0 + i, 'obj_id': oid, 'R': P[:, :3].copy(),
                                             't': P[:, 3:].copy(), 'score': float(ms_[q, k]), 'time': total / n})
                    continue
                r = recs[i, j]
                if r[14] != 1.0:
                    continue
                P = r[:12].reshape(3, 4)
                per_image[i].append({'scene_id': 0, 'im_id': b0 + i, 'obj_id': oid, 'R': P[:, :3].copy(), 't': P[:, 3:].copy(),
                                     'score': 0.0, 'time': total / n})    # score is always 0.0 in the reference (statistics.h:67)
        if args.vis:
            labels = out[model.PRED_OBJ_LABEL].cpu().numpy()
            for i in range(n):
                visualize(os.path.join(vis_dir, '{:06d}_grid.jpg'.format(b0 + i)), imgs_np[i], labels[i], per_image[i], K)
        for lst in per_image:
            results.extend(lst)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        results = [r for part in gathered for r in part]
        dist.destroy_process_group()
    if rank == 0 and args.save_estimates:
        os.makedirs(args.infer_dir, exist_ok=True)
        suffix = '_{}'.format(args.infer_name) if args.infer_name else ''
        path = os.path.join(args.infer_dir, 'estimated-poses{}.csv'.format(suffix))
        bop_io.save_bop_results(path, results)
        print('Saved {} pose estimates to: {}'.format(len(results), path))
    return results


if __name__ == '__main__':
    main()
