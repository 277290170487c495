"""Per-batch inference engine: the body of `process_image` (/root/reference/scripts/infer.py:348-554) for a
batch of images on one GPU -- model.predict -> establish_many_to_many -> per-object pose fitting -- with every
stage a hand-written sm_100a kernel behind the C ABI.  PyTorch supplies device memory, streams and (in
epos_b200/dist.py) the NCCL process group; nothing here computes on the host and there is no CPU fallback.

Stages can be switched off to run BASELINE config 2 (CNN only)."""
import numpy as np
import torch

from . import _lib, model

STAGES_CNN = 'cnn'
STAGES_FULL = 'cnn+corresp+fit'


class Engine:
    def __init__(self, weights, num_objs, num_frags, device, stages=STAGES_CNN, model_store=None, K=None,
                 fit_params=None, max_correspondences=4096, seed=0, min_obj_conf=0.1, min_frag_rel_conf=0.5,
                 model_options=None):
        self.dev = torch.device(device)
        self.net = model.EposNet(weights, num_objs, num_frags, self.dev, model_options=model_options)
        self.O, self.F = num_objs, num_frags
        self.stages = stages
        self.model_store = model_store
        self.K = K
        self.fit_params = fit_params
        self.max_corr = max_correspondences
        self.seed = seed
        self._fitter = None
        if stages == STAGES_FULL:
            from . import posefit
            self._fitter = posefit.BatchFitter(self.dev, num_objs, num_frags, model_store, K, fit_params,
                                               max_correspondences, seed, min_obj_conf=min_obj_conf,
                                               min_frag_rel_conf=min_frag_rel_conf)

    def run_device(self, images_dev):
        """images_dev [B,H,W,3] f32 CUDA.  Returns a dict of CUDA tensors: model.predict's outputs for 'cnn',
        plus 'poses' [B,O,16] f64 pose records for the full path."""
        out = self.net.predict(images_dev)
        if self._fitter is not None:
            out['poses'] = self._fitter.fit(out)
        return out

    def result_tensor(self, out):
        """The tensor a caller reads back per batch: pose records (full path) or the object label map."""
        return out['poses'] if 'poses' in out else out[model.PRED_OBJ_LABEL]

    def run_host(self, images_pinned, result_pinned=None):
        """End-to-end call with HOST buffers: H2D of the batch, the hot path, D2H of the result."""
        x = images_pinned.to(self.dev, non_blocking=True)
        out = self.run_device(x)
        r = self.result_tensor(out)
        if result_pinned is None:
            return r.cpu()
        result_pinned.copy_(r, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return result_pinned
