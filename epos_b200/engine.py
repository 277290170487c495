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
    """pipelined=True (full path only): the CNN of batch i+1 runs on the caller's stream while correspondence extraction
    and pose fitting of batch i run on a side stream.  Pose fitting is latency-bound (one CTA per (image, object)
    problem, a few long problems at the end) and leaves most SMs idle; the CNN kernels fill them -- the tcgen05 GEMM
    takes its tiles from a global counter, so it runs on whatever SMs are free.  Results are identical to the serial
    order (every problem has its own seed and workspace)."""

    def __init__(self, weights, num_objs, num_frags, device, stages=STAGES_CNN, model_store=None, K=None,
                 fit_params=None, max_correspondences=4096, seed=0, min_obj_conf=0.1, min_frag_rel_conf=0.5,
                 model_options=None, pipelined=False, post_fit=None, lazy_loc=True, multi_params=None, graphs=False):
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
        self.post_fit = post_fit                 # e.g. the NCCL all-gather of pose records (epos_b200/dist.py)
        # full path: pred_frag_loc is evaluated only at the surviving correspondences (model.EposNet.heads); the CNN-only
        # stage keeps the materialising mode = the output contract of model.predict
        self.lazy_loc = bool(lazy_loc) and stages == STAGES_FULL
        # with it, the fragment softmax (when not fused into the GEMM) runs only on the (pixel, object) pairs above min_obj_conf
        self.sparse_conf = float(min_obj_conf) if self.lazy_loc else None
        self.pipelined = bool(pipelined) and stages == STAGES_FULL
        self._side = torch.cuda.Stream(device=self.dev) if self.pipelined else None
        self._copy = torch.cuda.Stream(device=self.dev) if self.pipelined else None    # H2D of the next batch
        self._inflight = []                      # (maps kept alive, event after corresp) of batches still on the side stream
        self._last_fit = None                    # event after the most recent fit
        self._pending_host = None                # (pinned result, event) of the previous run_host call
        # graphs=True: the per-batch pipeline is captured once per input shape into CUDA graphs (two buffer sets: the CNN
        # of batch i+1 writes set (i+1) % 2 while correspondences / pose fitting of batch i read set i % 2) and replayed;
        # the host then issues ~4 launches per batch instead of ~170 kernel launches and ~200 allocations
        self.graphs = bool(graphs)
        self._gsets = {}                         # (B, H, W) -> [GraphSet, GraphSet]
        self._gturn = 0
        self.graph_launches = 0                  # kernel launches replayed through graphs (bench.py's gpu_launches)
        if stages == STAGES_FULL:
            from . import posefit
            self._fitter = posefit.BatchFitter(self.dev, num_objs, num_frags, model_store, K, fit_params,
                                               max_correspondences, seed, min_obj_conf=min_obj_conf,
                                               min_frag_rel_conf=min_frag_rel_conf, mparams=multi_params)

    def _fit(self, out, after_extract=None, K=None, num_instances=None):
        poses = self._fitter.fit(out, after_extract, K, num_instances).clone()   # the record buffer is reused by the next batch
        out['poses'] = poses
        if self._fitter.multi is not None:
            out['multi'] = self._fitter.multi                        # Progressive-X problems: every instance (BatchFitter.fit_maps)
        if self.post_fit is not None:
            out['poses_all'] = self.post_fit(poses)

    # ---- CUDA-graph mode -------------------------------------------------------------------------------------
    def _capture(self, shape, K):
        """Two buffer sets for input shape (B, H, W, 3): static input, CNN graph (main stream) and, for the full path, a
        post-processing graph (correspondences + pose fitting; side stream)."""
        lib = _lib.lib()
        B = shape[0]
        bi0 = 0
        if self._fitter is not None:
            self._fitter._prepare(B, K)
            bi0 = self._fitter.batch_index
        side = self._side if self._side is not None else torch.cuda.Stream(device=self.dev)
        self._side = side
        warm = torch.zeros(shape, dtype=torch.float32, device=self.dev)
        # eager warm-up on the capture stream: one-time initialisation (kernel attributes, scheduler ring, TMA encoder)
        cap = torch.cuda.Stream(device=self.dev)
        cap.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(cap):
            out = self.net.predict(warm, lazy_loc=self.lazy_loc, sparse_conf=self.sparse_conf)
            if self._fitter is not None:
                self._fitter.fit(out)
        cap.synchronize()
        sets = []
        pool = None
        for s in range(2):
            gs = {'x': torch.zeros(shape, dtype=torch.float32, device=self.dev)}
            g = torch.cuda.CUDAGraph()
            l0 = lib.epos_launch_count()
            with torch.cuda.graph(g, pool=pool, stream=cap, capture_error_mode='thread_local'):
                gs['out'] = self.net.predict(gs['x'], lazy_loc=self.lazy_loc, sparse_conf=self.sparse_conf)
            gs['cnn'] = g
            gs['cnn_launches'] = int(lib.epos_launch_count() - l0)
            # Serial engine: the two CNN graphs replay strictly one after the other, so they may share a private pool.
            # Pipelined engine: they must NOT -- with a shared pool the outputs of set 1 can be placed in blocks that
            # were intermediates of set 0's graph, and set 0's next replay would overwrite them while the side stream is
            # still post-processing set 1.
            pool = None if self.pipelined else g.pool()
            gs['post'] = None
            gs['post_launches'] = 0
            if self._fitter is not None:
                gp = torch.cuda.CUDAGraph()
                l0 = lib.epos_launch_count()
                with torch.cuda.graph(gp, stream=cap, capture_error_mode='thread_local'):
                    gs['poses'] = self._fitter.fit(gs['out']).clone()
                gs['post'] = gp
                gs['post_launches'] = int(lib.epos_launch_count() - l0)
            gs['ev_cnn'] = torch.cuda.Event()
            gs['ev_post'] = torch.cuda.Event()
            sets.append(gs)
        # the warm-up and the captures advanced the host-side counter only; the device-side batch counter restarts here
        if self._fitter is not None:
            self._fitter.batch_index = bi0
            with torch.cuda.stream(cap):
                self._fitter._batch_dev.fill_(bi0)
        torch.cuda.current_stream(self.dev).wait_stream(cap)
        return sets

    def _run_graph(self, images, K=None, from_host=False):
        """One batch through the captured graphs.  images: device tensor, or pinned host tensor with from_host=True
        (copied straight into the static input on the copy stream)."""
        shape = tuple(images.shape)
        key = shape[:3]
        if key not in self._gsets:
            self._gsets[key] = self._capture(shape, K)
        gs = self._gsets[key][self._gturn % 2]
        self._gturn += 1
        main = torch.cuda.current_stream(self.dev)
        main.wait_event(gs['ev_post'])           # batch i-2 has finished reading this set's head maps / poses
        if from_host and self._copy is not None:
            with torch.cuda.stream(self._copy):
                self._copy.wait_event(gs['ev_cnn'])          # the CNN of batch i-2 has consumed the static input
                self._copy.wait_event(gs['ev_post'])
                gs['x'].copy_(images, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self._copy)
            main.wait_event(ev_in)
        else:
            gs['x'].copy_(images, non_blocking=True)
        gs['cnn'].replay()
        gs['ev_cnn'].record(main)
        self.graph_launches += gs['cnn_launches']
        out = dict(gs['out'])
        if gs['post'] is None:
            gs['ev_post'].record(main)
            return out
        post_stream = self._side if self.pipelined else main
        with torch.cuda.stream(post_stream):
            post_stream.wait_event(gs['ev_cnn'])
            # intrinsics of this batch (outside the graph, on the stream of the fit: ordered behind the previous batch's fit)
            self._fitter._prepare(shape[0], K)
            gs['post'].replay()
            self._fitter.batch_index += 1
            self.graph_launches += gs['post_launches']
            out['poses'] = gs['poses']
            if self.post_fit is not None:
                out['poses_all'] = self.post_fit(gs['poses'])
            gs['ev_post'].record(post_stream)
        self._last_fit = gs['ev_post']
        out['ready'] = gs['ev_post']
        return out

    def launch_count(self):
        """Kernels of this library launched so far, including those replayed through CUDA graphs."""
        return int(_lib.lib().epos_launch_count()) + self.graph_launches

    def run_device(self, images_dev, K=None, num_instances=None):
        """images_dev [B,H,W,3] f32 CUDA; K = this batch's camera intrinsics, [3,3] or [B,3,3] (default: the K given
        at construction -- the reference reads K per image, scripts/infer.py:376-377).  Returns a dict of CUDA tensors:
        model.predict's outputs for 'cnn', plus 'poses' [B,O,16] f64 pose records for the full path (pipelined: valid
        after join() / out['ready']).  num_instances [B,J]: instance bound per (image, object slot) as in
        scripts/infer.py:462-468 (0 = skip, 1 = GC-RANSAC, 2.. = Progressive-X, -1 = all); out['multi'] then holds every
        instance of the Progressive-X problems."""
        if self.graphs and num_instances is None:
            return self._run_graph(images_dev, K)
        out = self.net.predict(images_dev, lazy_loc=self.lazy_loc, sparse_conf=self.sparse_conf)
        if self._fitter is None:
            return out
        if not self.pipelined:
            self._fit(out, K=K, num_instances=num_instances)
            return out
        main = torch.cuda.current_stream(self.dev)
        done_cnn = torch.cuda.Event()
        done_cnn.record(main)
        # keep at most two batches of head maps alive; the older one is dropped once its correspondences are extracted
        while len(self._inflight) >= 2:
            _, ev = self._inflight.pop(0)
            ev.synchronize()
        with torch.cuda.stream(self._side):
            self._side.wait_event(done_cnn)
            maps = [out[k] for k in (model.PRED_OBJ_CONF, model.PRED_FRAG_CONF, model.PRED_FRAG_LOC) if k in out]
            if model.LAZY_FRAG_LOC in out:
                maps.append(out[model.LAZY_FRAG_LOC][0])           # decoder features read by the lazy localisation head
            for t in maps:
                t.record_stream(self._side)
            ev_maps = torch.cuda.Event()
            self._fit(out, after_extract=lambda: ev_maps.record(self._side), K=K, num_instances=num_instances)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._inflight.append((maps, ev_maps))
        self._last_fit = ev
        out['ready'] = ev
        return out

    def join(self):
        """Makes the caller's stream wait for everything queued on the side stream."""
        if self._last_fit is not None:
            torch.cuda.current_stream(self.dev).wait_event(self._last_fit)

    def result_tensor(self, out):
        """The tensor a caller reads back per batch: pose records (full path) or the object label map."""
        return out['poses'] if 'poses' in out else out[model.PRED_OBJ_LABEL]

    def run_host(self, images_pinned, result_pinned=None, K=None):
        """End-to-end call with HOST buffers: H2D of the batch, the hot path, D2H of the result.
        Serial engine: returns this batch's result.  Pipelined engine: the D2H of this batch is queued behind its pose
        fitting on the side stream and the call returns the PREVIOUS batch's result (None on the first call); flush()
        returns the last one -- the host always holds batch i while the GPU works on batch i+1."""
        if self.graphs:
            out = self._run_graph(images_pinned, K, from_host=True)
            r = self.result_tensor(out)
            if not self.pipelined:
                if result_pinned is None:
                    return r.cpu()
                result_pinned.copy_(r, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return result_pinned
            prev = self.flush()
            buf = torch.empty(r.shape, dtype=r.dtype).pin_memory() if result_pinned is None else result_pinned
            with torch.cuda.stream(self._side):
                buf.copy_(r, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._side)
                out['ready'].record(self._side)      # the next user of this buffer set also waits for the read-back
            self._pending_host = (buf, ev)
            return prev
        if self.pipelined:
            # the copy engine brings batch i+1 in while the SMs are still busy with batch i
            main = torch.cuda.current_stream(self.dev)
            with torch.cuda.stream(self._copy):
                x = images_pinned.to(self.dev, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self._copy)
            main.wait_event(ev_in)
            x.record_stream(main)
        else:
            x = images_pinned.to(self.dev, non_blocking=True)
        out = self.run_device(x, K=K)
        r = self.result_tensor(out)
        if not self.pipelined:
            if result_pinned is None:
                return r.cpu()
            result_pinned.copy_(r, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return result_pinned
        prev = self.flush()
        buf = torch.empty(r.shape, dtype=r.dtype).pin_memory() if result_pinned is None else result_pinned
        with torch.cuda.stream(self._side):
            buf.copy_(r, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._pending_host = (buf, ev)
        return prev

    def flush(self):
        """Pipelined engine: waits for and returns the host result of the last run_host call (None if none pending)."""
        if self._pending_host is None:
            return None
        buf, ev = self._pending_host
        self._pending_host = None
        ev.synchronize()
        return buf
