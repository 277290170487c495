#!/usr/bin/env python
"""Entry point with the flags of /root/reference/scripts/infer.py (:37-146) for the B200 path.

The reference reads TFRecords and a TF checkpoint; this build has neither TensorFlow nor the BOP datasets, so images
are synthetic (--synthetic, the default and only mode so far) and weights are random-init or an .npz keyed by the TF
variable names (--weights).  Per-image flow, timing keys and the BOP CSV row format follow process_image
(infer.py:348-554), main (:712-760) and bop_toolkit inout.save_bop_results (inout.py:265-294).

  python scripts/infer.py --num_images 16 --batch_size 8 --num_objs 21 --num_frags 64 [--world_size N via torchrun]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    # flags of the reference (same names and defaults)
    ap.add_argument('--model', default='synthetic')
    ap.add_argument('--task_type', default='localization', choices=['localization', 'detection'])
    ap.add_argument('--infer_name', default=None)
    ap.add_argument('--fitting_method', default='progressive_x', choices=['progressive_x'])
    ap.add_argument('--inlier_thresh', type=float, default=4.0)
    ap.add_argument('--neighbour_max_dist', type=float, default=20.0)
    ap.add_argument('--min_hypothesis_quality', type=float, default=0.5)
    ap.add_argument('--required_progx_confidence', type=float, default=0.5)
    ap.add_argument('--required_ransac_confidence', type=float, default=1.0)
    ap.add_argument('--min_triangle_area', type=float, default=0.0)
    ap.add_argument('--use_prosac', action='store_true')
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
    ap.add_argument('--save_estimates', action='store_true', default=True)
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
    return ap.parse_args()


def save_bop_results(path, results):
    """CSV rows of inout.save_bop_results(version='bop19')."""
    lines = ['scene_id,im_id,obj_id,score,R,t,time']
    for r in results:
        lines.append('{scene_id},{im_id},{obj_id},{score},{R},{t},{time}'.format(
            scene_id=r['scene_id'], im_id=r['im_id'], obj_id=r['obj_id'], score=r['score'],
            R=' '.join(map(str, r['R'].flatten().tolist())), t=' '.join(map(str, r['t'].flatten().tolist())),
            time=r.get('time', -1)))
    with open(path, 'w') as f:
        f.write('\n'.join(lines))


def main():
    args = parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from epos_b200 import dist as edist, engine, model, posefit, synthetic, weights as W
    if args.required_ransac_confidence != 1.0:
        raise SystemExit('required_ransac_confidence must be 1.0 in this build')
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
    eng = engine.Engine(w, O, F, dev, stages=engine.STAGES_FULL, model_store=store, K=K, fit_params=params,
                        max_correspondences=args.max_correspondences, seed=args.seed,
                        min_obj_conf=args.corr_min_obj_conf, min_frag_rel_conf=args.corr_min_frag_rel_conf,
                        model_options=model.ModelOptions(W.head_channels(O, F), model_variant=args.model_variant))
    lo, hi = edist.shard_range(args.num_images, world, rank)
    results = []
    ids = store.dp_model['obj_ids']
    for b0 in range(lo, hi, args.batch_size):
        n = min(args.batch_size, hi - b0)
        imgs = torch.from_numpy(np.concatenate([W.synthetic_images(1, seed=10000 + i) for i in range(b0, b0 + n)])).pin_memory()
        torch.cuda.synchronize()
        t0 = time.time()
        recs = eng.run_host(imgs).numpy()                     # [n, J, 16]
        total = time.time() - t0
        print('Images: {}-{}, total time: {:.3f} s ({:.1f} images/s)'.format(b0, b0 + n - 1, total, n / total), flush=True)
        for i in range(n):
            for j, oid in enumerate(ids):
                r = recs[i, j]
                if r[14] != 1.0:
                    continue
                P = r[:12].reshape(3, 4)
                results.append({'scene_id': 0, 'im_id': b0 + i, 'obj_id': oid, 'R': P[:, :3].copy(), 't': P[:, 3:].copy(),
                                'score': 0.0, 'time': total / n})    # score is always 0.0 in the reference (statistics.h:67)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        results = [r for part in gathered for r in part]
        dist.destroy_process_group()
    if rank == 0 and args.save_estimates:
        os.makedirs(args.infer_dir, exist_ok=True)
        suffix = '_{}'.format(args.infer_name) if args.infer_name else ''
        path = os.path.join(args.infer_dir, 'estimated-poses{}.csv'.format(suffix))
        save_bop_results(path, results)
        print('Saved {} pose estimates to: {}'.format(len(results), path))


if __name__ == '__main__':
    main()
