"""Pose fitting on the GPU: drop-in for pyprogressivex.find6DPoses
(/root/reference/external/progressive-x/src/pyprogressivex/src/bindings.cpp:9-155), single-instance branch, and the
batched fitter the engine uses after model.predict.  All arithmetic is in csrc/posefit.cu behind epos_fit_poses();
there is no CPU fallback."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import corresp as _corresp


def default_params(**kw):
    p = _lib.FitParams()
    _lib.lib().epos_fit_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class PoseFitter:
    """P independent single-instance problems per call (one CTA each)."""

    def __init__(self, device, max_problems, params=None):
        self.lib = _lib.lib()
        self.dev = torch.device(device)
        self.params = params or default_params()
        self.P = max_problems
        nbytes = self.lib.epos_fit_workspace_bytes(max_problems, self.lib.epos_fit_max_points(), C.byref(self.params))
        self.workspace = torch.empty((nbytes + 256,), dtype=torch.uint8, device=self.dev)
        self._ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256

    def fit(self, coord_2d, coord_3d, offsets, counts, K, seeds, poses=None, labeling=None):
        """coord_2d [R,2] f64, coord_3d [R,3] f64, offsets/counts [P] i32, K [P,3,3] f64, seeds [P] i64 (device).
        Returns (poses [P,16] f64, labeling [R] i32)."""
        P = offsets.numel()
        assert P <= self.P
        if poses is None:
            poses = torch.empty((P, 16), dtype=torch.float64, device=self.dev)
        if labeling is None:
            labeling = torch.empty((coord_2d.shape[0],), dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.epos_fit_poses(coord_2d.data_ptr(), coord_3d.data_ptr(), offsets.data_ptr(), counts.data_ptr(),
                                           P, K.data_ptr(), seeds.data_ptr(), C.byref(self.params), poses.data_ptr(),
                                           labeling.data_ptr(), self._ws_ptr, self.workspace.numel() - 256,
                                           _lib.stream_ptr()), 'epos_fit_poses')
        return poses, labeling

    def debug_state(self, P):
        """Per-problem counters of the last fit (synchronises): [P, EPOS_FIT_DEBUG_COLS] i64, columns as in
        include/epos_b200.h."""
        out = np.zeros((P, 20), np.int64)
        _lib.check(self.lib.epos_fit_debug_state(self._ws_ptr, P, out.ctypes.data), 'epos_fit_debug_state')
        return out

    def algorithmic_bytes(self, P):
        """SURVEY.md 8d RANSAC figure for the last fit: every model scored over a problem's N points reads
        N x 40 B (u_n, v_n, x, y, z in f64) + N x 8 B (pixel id); points the main loop's early-out never visited
        (column 19) are not counted."""
        d = self.debug_state(P)
        return float((d[:, 0] * (d[:, 16] + d[:, 17] + d[:, 18]) - d[:, 19]).sum()) * 48.0


def multi_params(max_model_number_for_pearl=5, min_point_number=6, confidence=0.5, max_tanimoto_similarity=0.9):
    return _lib.MultiParams(int(max_model_number_for_pearl), int(min_point_number), float(confidence),
                            float(max_tanimoto_similarity))


class MultiPoseFitter:
    """P independent multi-instance (Progressive-X + PEARL) problems per call, one persistent CTA each
    (epos_fit_poses_multi).  max_models [P]: the instance bound of each problem, 2 .. max_model_number_for_pearl."""

    def __init__(self, device, max_problems, params=None, mparams=None):
        self.lib = _lib.lib()
        self.dev = torch.device(device)
        self.params = params or default_params()
        self.mparams = mparams or multi_params()
        self.P = max_problems
        self.MAXI = self.lib.epos_fit_max_instances()
        nbytes = self.lib.epos_fit_multi_workspace_bytes(max_problems)
        self.workspace = torch.empty((nbytes + 256,), dtype=torch.uint8, device=self.dev)
        self._ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256

    def fit(self, coord_2d, coord_3d, offsets, counts, K, seeds, max_models):
        """Device tensors as PoseFitter.fit + max_models [P] i32.  Returns (records [P,16], labeling [R] i32,
        multi_poses [P,MAXI,12], multi_scores [P,MAXI], multi_counts [P] i32)."""
        P = offsets.numel()
        assert P <= self.P and max_models.numel() == P and max_models.dtype == torch.int32
        poses = torch.zeros((P, 16), dtype=torch.float64, device=self.dev)
        labeling = torch.zeros((coord_2d.shape[0],), dtype=torch.int32, device=self.dev)
        mposes = torch.zeros((P, self.MAXI, 12), dtype=torch.float64, device=self.dev)
        mscores = torch.zeros((P, self.MAXI), dtype=torch.float64, device=self.dev)
        mcounts = torch.zeros((P,), dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.epos_fit_poses_multi(
            coord_2d.data_ptr(), coord_3d.data_ptr(), offsets.data_ptr(), counts.data_ptr(), P, K.data_ptr(),
            seeds.data_ptr(), C.byref(self.params), C.byref(self.mparams), max_models.data_ptr(), poses.data_ptr(),
            labeling.data_ptr(), mposes.data_ptr(), mscores.data_ptr(), mcounts.data_ptr(), self._ws_ptr,
            self.workspace.numel() - 256, _lib.stream_ptr()), 'epos_fit_poses_multi')
        return poses, labeling, mposes, mscores, mcounts

    def debug_state(self, P):
        out = np.zeros((P, 8), np.int64)
        _lib.check(self.lib.epos_fit_multi_debug_state(self._ws_ptr, P, out.ctypes.data), 'epos_fit_multi_debug_state')
        return out


class BatchFitter:
    """model.predict outputs -> correspondences -> poses for a whole batch: [B, J, 16] pose records on the device
    (record layout: include/epos_b200.h EPOS_POSE_RECORD_DOUBLES)."""

    def __init__(self, device, num_objs, num_frags, model_store, K, params=None, max_correspondences=4096, seed=0,
                 obj_ids=None, output_scale=0.25, min_obj_conf=0.1, min_frag_rel_conf=0.5, mparams=None):
        self.dev = torch.device(device)
        self.params = params or default_params()
        self.mparams = mparams or multi_params()
        self._multi_fitter = None
        self.multi = None                 # result of the Progressive-X problems of the last batch (see fit_maps)
        nmax = _lib.lib().epos_fit_max_points()
        if max_correspondences is None or max_correspondences > nmax:
            raise ValueError('this build keeps the point set in shared memory: max_correspondences must be <= %d '
                             '(the reference default is None = unbounded, infer.py:112-114)' % nmax)
        self.extract = _corresp.CorrespExtractor(self.dev, num_objs, num_frags, model_store, obj_ids=obj_ids,
                                                 output_scale=output_scale, min_obj_conf=min_obj_conf,
                                                 min_frag_rel_conf=min_frag_rel_conf, cap=max_correspondences,
                                                 max_correspondences=max_correspondences)
        self.J = len(self.extract.obj_ids_list)
        self.K = None if K is None else np.asarray(K, np.float64)     # default intrinsics; a call may pass its own
        self.seed = int(seed)
        self._fitter = None
        self._Kdev = None
        self._Khost = None
        self._P = 0
        self.batch_index = 0
        self._batch_dev = None

    def _prepare(self, B, K=None):
        """Buffers for B images and the intrinsics of this batch.  The reference reads K per image
        (samples[common.K][0], scripts/infer.py:376-377; BOP sets such as T-LESS have per-image intrinsics), so K is
        [3,3] (all images) or [B,3,3]; the device copy is refreshed only when the values change."""
        P = B * self.J
        if self._fitter is None or self._fitter.P < P:
            self._fitter = PoseFitter(self.dev, P, self.params)
        if self._P != P:
            self._P = P
            self._Kdev = torch.empty((P, 9), dtype=torch.float64, device=self.dev)
            self._Khost = None
            self._poses = torch.empty((P, 16), dtype=torch.float64, device=self.dev)
            self._labeling = torch.empty((P * self.extract.cap,), dtype=torch.int32, device=self.dev)
            self._base = torch.arange(P, dtype=torch.int64, device=self.dev) + (self.seed << 32)
        if self._batch_dev is None:
            # the batch counter of the stream keys lives on the device so that a captured CUDA graph advances it itself
            self._batch_dev = torch.full((1,), self.batch_index, dtype=torch.int64, device=self.dev)
        K = self.K if K is None else np.asarray(K, np.float64)
        if K is None:
            raise ValueError('camera intrinsics K are required ([3,3] or [B,3,3])')
        if K.shape == (3, 3):
            K = np.broadcast_to(K, (B, 3, 3))
        if K.shape != (B, 3, 3):
            raise ValueError('K should be [3,3] or [%d,3,3], got %s' % (B, K.shape))
        if self._Khost is None or not np.array_equal(self._Khost, K):
            self._Khost = np.array(K, np.float64)
            rows = np.ascontiguousarray(np.repeat(self._Khost.reshape(B, 9), self.J, axis=0))
            # asynchronous, ordered on the CURRENT stream behind the fit of the previous batch (which may still read the old
            # values); the pinned staging buffer is kept alive until the next change
            self._Kpin = torch.from_numpy(rows).pin_memory()
            self._Kdev.copy_(self._Kpin, non_blocking=True)

    def seeds_for(self, B):
        """Stream key of problem (image b, slot j) in batch n: seed * 2^32 + n * P + b * J + j.  Advances the device-side
        batch counter (graph-capturable: no host value is baked in)."""
        seeds = self._base + self._batch_dev * (B * self.J)
        self._batch_dev += 1
        return seeds

    def fit_maps(self, obj_conf, frag_conf, frag_loc, after_extract=None, K=None, lazy_loc=None, num_instances=None):
        """num_instances [B, J] (host ints; None = 1 everywhere) is the `num_instances` of scripts/infer.py:462-468 per
        (image, object slot): 0 = object not annotated (skipped), 1 = GC-RANSAC + final LM, 2 .. = Progressive-X with
        that instance bound, -1 = all instances (DETECTION).  Records [B, J, 16]; for Progressive-X problems the record
        holds the first instance and record[14] the number of instances, all of them are in self.multi =
        {'index': [(b, j), ...], 'poses' [Pm, MAXI, 12], 'scores' [Pm, MAXI], 'counts' [Pm], 'labeling'} (device)."""
        B = obj_conf.shape[0]
        self._prepare(B, K)
        bc = self.extract(obj_conf, frag_conf, frag_loc, lazy_loc=lazy_loc)
        self.corr = bc
        if after_extract is not None:
            after_extract()                       # the head maps are no longer needed from here on
        seeds = self.seeds_for(B)
        self.batch_index += 1
        counts = bc.counts
        self.multi = None
        midx = None
        if num_instances is not None:
            ni = np.ascontiguousarray(num_instances, np.int32).reshape(-1)
            if ni.size != B * self.J:
                raise ValueError('num_instances should be [B, J] = [%d, %d]' % (B, self.J))
            if ((ni < -1)).any():
                raise ValueError('num_instances entries should be -1, 0 or positive')
            if (ni != 1).any():
                nid = torch.from_numpy(ni).to(self.dev)
                counts = torch.where(nid == 1, counts, torch.zeros_like(counts))
                midx = np.nonzero((ni != 1) & (ni != 0))[0]
        poses, lab = self._fitter.fit(bc.coord_2d.view(-1, 2), bc.coord_3d.view(-1, 3), bc.offsets, counts,
                                      self._Kdev, seeds, self._poses, self._labeling)
        if midx is not None and midx.size:
            if self._multi_fitter is None or self._multi_fitter.P < midx.size:
                self._multi_fitter = MultiPoseFitter(self.dev, max(int(midx.size), 4), self.params, self.mparams)
            it = torch.from_numpy(midx.astype(np.int64)).to(self.dev)
            rec, mlab, mposes, mscores, mcounts = self._multi_fitter.fit(
                bc.coord_2d.view(-1, 2), bc.coord_3d.view(-1, 3), bc.offsets[it].contiguous(), bc.counts[it].contiguous(),
                self._Kdev[it].contiguous(), seeds[it].contiguous(), nid[it].contiguous())
            poses.index_copy_(0, it, rec)
            self.multi = {'index': [(int(i) // self.J, int(i) % self.J) for i in midx], 'poses': mposes, 'scores': mscores,
                          'counts': mcounts, 'labeling': mlab}
        return poses.view(B, self.J, 16)

    def fit(self, predictions, after_extract=None, K=None, num_instances=None):
        from . import model
        return self.fit_maps(predictions[model.PRED_OBJ_CONF], predictions[model.PRED_FRAG_CONF],
                             predictions.get(model.PRED_FRAG_LOC), after_extract, K,
                             lazy_loc=predictions.get(model.LAZY_FRAG_LOC), num_instances=num_instances)


def find6DPoses(x1y1, x2y2z2, K, threshold=4.0, max_model_number=-1, conf=0.5, proposal_engine_conf=1.0,
                spatial_coherence_weight=0.1, neighborhood_ball_radius=20.0, max_tanimoto_similarity=0.9,
                scaling_from_millimeters=0.1, min_triangle_area=100.0, min_coverage=0.5, max_iters=400,
                min_point_number=6, use_prosac=False, max_model_number_for_optimization=3,
                apply_numerical_optimization=True, log=False, seed=0, device='cuda:0', max_neighbors=5):
    """Same name / keyword arguments / return triple as pyprogressivex.find6DPoses (bindings.cpp:9-28,117,133-152):
    (poses [3M,4] f64, labeling [N] i32, scores [M] f64) as numpy arrays.  Errors: ValueError for malformed shapes
    (bindings.cpp:30-58).  max_model_number == 1: GC-RANSAC + final LM; 2 .. max_model_number_for_optimization:
    Progressive-X with PEARL (one pose per instance, labeling = instance index, scores = instance support); more, or -1:
    sequential propose-and-remove fitting (spedUpFitting; labeling and scores zero as in the reference; -1 is bounded,
    see include/epos_b200.h).  `seed` selects the RANSAC stream (the reference seeds from std::random_device)."""
    x1y1 = np.ascontiguousarray(x1y1, np.float64)
    x2y2z2 = np.ascontiguousarray(x2y2z2, np.float64)
    K = np.ascontiguousarray(K, np.float64)
    if x1y1.ndim != 2 or x1y1.shape[1] != 2:
        raise ValueError('x1y1 should be an array with dims [n,2], n>=3')
    n = x1y1.shape[0]
    if x2y2z2.ndim != 2 or x2y2z2.shape[1] != 3 or x2y2z2.shape[0] != n or n < 3:
        raise ValueError('x2y2z2 should be an array with dims [n,3], n>=3, same n as x1y1')
    if K.shape != (3, 3):
        raise ValueError('K should be an array with dims [3,3]')
    if proposal_engine_conf != 1.0:
        raise NotImplementedError('proposal_engine_conf must be 1.0 (scripts/infer.py:90)')
    dev = torch.device(device)
    if max_model_number != 1:
        # Progressive-X (progressivex_python.cpp:136-221).  PEARL mode: 2 .. max_model_number_for_optimization instances.
        if max_model_number == 0 or max_model_number < -1:
            raise ValueError('max_model_number should be -1 or positive')
        if max_model_number_for_optimization > 5:
            raise ValueError('max_model_number_for_optimization must be <= 5 (infer.py max_model_number_for_pearl default)')
        if n > _lib.lib().epos_fit_max_points():
            raise ValueError('more than %d correspondences' % _lib.lib().epos_fit_max_points())
        p = default_params(threshold=threshold, spatial_coherence_weight=spatial_coherence_weight,
                           neighborhood_ball_radius=neighborhood_ball_radius,
                           scaling_from_millimeters=scaling_from_millimeters, min_triangle_area=min_triangle_area,
                           min_coverage=min_coverage, max_iters=max_iters, max_neighbors=max_neighbors,
                           apply_numerical_optimization=0)
        mp = multi_params(max_model_number_for_optimization, min_point_number, conf, max_tanimoto_similarity)
        f = MultiPoseFitter(dev, 1, p, mp)
        rec, lab, mposes, mscores, mcounts = f.fit(
            torch.from_numpy(x1y1).to(dev), torch.from_numpy(x2y2z2).to(dev), torch.zeros(1, dtype=torch.int32, device=dev),
            torch.tensor([n], dtype=torch.int32, device=dev), torch.from_numpy(K.reshape(1, 3, 3)).to(dev),
            torch.tensor([seed], dtype=torch.int64, device=dev), torch.tensor([max_model_number], dtype=torch.int32, device=dev))
        m = int(mcounts.cpu().numpy()[0])
        if m < 0:
            raise RuntimeError('epos_fit_poses_multi rejected max_model_number=%d' % max_model_number)
        find6DPoses.last_record = rec.cpu().numpy()[0]
        find6DPoses.last_multi_state = f.debug_state(1)[0]
        return (mposes.cpu().numpy()[0, :m].reshape(3 * m, 4).copy(), lab.cpu().numpy().astype(np.int32),
                mscores.cpu().numpy()[0, :m].copy())
    p = default_params(threshold=threshold, spatial_coherence_weight=spatial_coherence_weight,
                       neighborhood_ball_radius=neighborhood_ball_radius,
                       scaling_from_millimeters=scaling_from_millimeters, min_triangle_area=min_triangle_area,
                       min_coverage=min_coverage, max_iters=max_iters, max_neighbors=max_neighbors,
                       apply_numerical_optimization=int(apply_numerical_optimization))
    if n > _lib.lib().epos_fit_max_points():
        raise ValueError('more than %d correspondences' % _lib.lib().epos_fit_max_points())
    f = PoseFitter(dev, 1, p)
    poses, lab = f.fit(torch.from_numpy(x1y1).to(dev), torch.from_numpy(x2y2z2).to(dev),
                       torch.zeros(1, dtype=torch.int32, device=dev), torch.tensor([n], dtype=torch.int32, device=dev),
                       torch.from_numpy(K.reshape(1, 3, 3)).to(dev), torch.tensor([seed], dtype=torch.int64, device=dev))
    rec = poses.cpu().numpy()[0]
    labeling = lab.cpu().numpy().astype(np.int32)
    find6DPoses.last_record = rec
    if rec[14] != 1.0:
        return np.zeros((0, 4)), labeling, np.zeros((0,))
    return rec[:12].reshape(3, 4).copy(), labeling, np.zeros((1,))
