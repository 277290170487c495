"""ORACLE (test infrastructure, never imported by the product): numpy restatement of
`establish_many_to_many` (/root/reference/epos_lib/corresp.py:9-101) and of the top-K selection the
inference script applies afterwards (/root/reference/scripts/infer.py:425-440).

Semantics kept on purpose:
  * object channel index is `obj_id` in obj_confs (0 = background) but `obj_id - 1` in the
    fragment tensors (corresp.py:46,60,76);
  * 2D coordinates are scale * (index + 0.5) in float64 (misc.py:14-26);
  * the local 3D offset is multiplied by the fragment size in FLOAT32 before being added to
    the float64 fragment centre (corresp.py:76-78);
  * output order: row-major pixel, then fragment id.
"""
import numpy as np


def convert_px_indices_to_im_coords(px_indices, scale):
    return scale * (px_indices.astype(np.float64) + 0.5)


def establish_many_to_many(obj_confs, frag_confs, frag_coords, gt_obj_ids, obj_ids, frag_centers,
                           frag_sizes, output_scale, min_obj_conf, min_frag_rel_conf,
                           only_annotated_objs=True):
    corresp = {}
    for obj_id in obj_ids:
        if only_annotated_objs and obj_id not in gt_obj_ids:
            continue
        obj_conf = obj_confs[:, :, obj_id]
        obj_mask = obj_conf > min_obj_conf
        if not np.any(obj_mask):
            continue
        yx = np.stack(np.nonzero(obj_mask), axis=0).T
        im_coords = convert_px_indices_to_im_coords(np.flip(yx, axis=1), 1.0 / output_scale)
        frag_conf_masked = frag_confs[obj_mask][:, obj_id - 1, :]
        frag_conf_max = np.max(frag_conf_masked, axis=1, keepdims=True)
        frag_mask = frag_conf_masked > (frag_conf_max * min_frag_rel_conf)
        frag_inds = np.stack(np.nonzero(frag_mask), axis=0).T
        corr_2d = im_coords[frag_inds[:, 0]]
        corr_3d = np.array(frag_centers[obj_id][frag_inds[:, 1]], dtype=np.float64)
        frag_scales = np.expand_dims(frag_sizes[obj_id][frag_inds[:, 1]], 1)
        corr_3d_local = frag_coords[obj_mask][:, obj_id - 1, :, :][frag_mask]      # float32 copy
        corr_3d_local *= frag_scales                                                 # f32 in-place multiply
        corr_3d += corr_3d_local
        conf_obj = obj_conf[obj_mask][frag_inds[:, 0]]
        conf_frag = frag_conf_masked[frag_mask]
        corresp[obj_id] = {
            'px_id': frag_inds[:, 0], 'frag_id': frag_inds[:, 1],
            'coord_2d': corr_2d, 'coord_3d': corr_3d,
            'conf': conf_obj * conf_frag, 'conf_obj': conf_obj, 'conf_frag': conf_frag,
            # linear output-map pixel index of every correspondence (not in the reference dict;
            # convenience for comparing with the device path)
            'pixel': (yx[:, 0] * obj_confs.shape[1] + yx[:, 1])[frag_inds[:, 0]],
        }
    return corresp


def select_top_k(obj_corr, max_correspondences):
    """infer.py:431-440 with use_prosac=False: keep the `max_correspondences` most confident rows
    in descending confidence order (np.argsort(conf)[::-1], i.e. ties broken by DESCENDING index)."""
    n = obj_corr['coord_2d'].shape[0]
    if max_correspondences is None or n <= max_correspondences:
        return obj_corr
    keep = np.argsort(obj_corr['conf'], kind='stable')[::-1][:max_correspondences]
    return {k: v[keep] for k, v in obj_corr.items()}
