"""ORACLE (test infrastructure): the post-processing half of process_image (/root/reference/scripts/infer.py:396-503)
on the CPU -- establish_many_to_many -> top-K by confidence -> per-object find6DPoses (single instance) -- chained from
the oracle pieces.  Used by tests and by bench.py's cpu_baseline / --impl reference legs only."""
import numpy as np

from . import corresp as ocorr
from . import posefit as opose


class PostProcess:
    def __init__(self, num_objs, num_frags, seed=0, model_store=None, K=None, max_correspondences=4096,
                 min_obj_conf=0.1, min_frag_rel_conf=0.5, output_scale=0.25, fit_kwargs=None):
        if model_store is None:
            from epos_b200 import synthetic      # scene description only (no compute): shared with the product
            model_store = synthetic.model_store(num_objs, num_frags)
            K = synthetic.default_K() if K is None else K
        self.store, self.K = model_store, np.asarray(K, np.float64)
        self.O, self.F = num_objs, num_frags
        self.seed = int(seed)
        self.max_corr = max_correspondences
        self.min_obj_conf, self.min_frag_rel_conf, self.output_scale = min_obj_conf, min_frag_rel_conf, output_scale
        self.fit_kwargs = dict(threshold=4.0, min_triangle_area=0.0, max_iters=400)
        self.fit_kwargs.update(fit_kwargs or {})

    def corresp(self, out, b=0):
        ids = self.store.dp_model['obj_ids']
        c = ocorr.establish_many_to_many(out['pred_obj_conf'][b], out['pred_frag_conf'][b], out['pred_frag_loc'][b],
                                         ids, ids, self.store.frag_centers, self.store.frag_sizes, self.output_scale,
                                         self.min_obj_conf, self.min_frag_rel_conf, only_annotated_objs=False)
        return {oid: ocorr.select_top_k(d, self.max_corr) for oid, d in c.items()}

    def seed_for(self, image_index, slot, n_slots, images_per_batch=1, batch_index=None):
        """Same stream key as epos_b200.posefit.BatchFitter.seeds_for."""
        if batch_index is None:
            batch_index, b = divmod(image_index, images_per_batch)
        else:
            b = image_index
        return (self.seed << 32) + batch_index * (images_per_batch * n_slots) + b * n_slots + slot

    def fit(self, corr, image_index=0, images_per_batch=1, batch_index=None):
        """-> {obj_id: record[16]} with the layout of EPOS_POSE_RECORD_DOUBLES."""
        ids = self.store.dp_model['obj_ids']
        recs = {}
        for j, oid in enumerate(ids):
            rec = np.zeros(16)
            d = corr.get(oid)
            if d is not None and d['coord_2d'].shape[0] >= 6:          # infer.py:417-422
                poses, lab, _, st = opose.find6DPoses(
                    np.ascontiguousarray(d['coord_2d'], np.float64), np.ascontiguousarray(d['coord_3d'], np.float64),
                    self.K, max_model_number=1, seed=self.seed_for(image_index, j, len(ids), images_per_batch, batch_index),
                    return_stats=True, **self.fit_kwargs)
                rec[13], rec[15] = st['iterations'], st['graph_cuts']
                if poses.shape[0] == 3:
                    rec[:12] = poses.ravel()
                    rec[12], rec[14] = lab.sum(), 1.0
                rec_lab = lab
            recs[oid] = rec
        return recs
