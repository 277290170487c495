"""Drop-in for /root/reference/epos_lib/corresp.py:9-101 (`establish_many_to_many`) on the GPU, plus the batched form
used by the engine.  All arithmetic is in csrc/corresp.cu behind epos_corresp(); there is no CPU fallback."""
import numpy as np
import torch

from . import _lib


class BatchCorresp:
    """Correspondences for every (image, object slot) of a batch, kept on the device.

    Fields (device tensors): coord_2d [S,cap,2] f64, coord_3d [S,cap,3] f64, conf/conf_obj/conf_frag [S,cap] f32,
    px/frag [S,cap] i32, counts [S] i32, totals [S] i32 with S = B*J segments (segment = b*J + j)."""

    def __init__(self, device, B, J, cap, h, w):
        dev = torch.device(device)
        S = B * J
        self.B, self.J, self.cap, self.h, self.w = B, J, cap, h, w
        self.coord_2d = torch.empty((S, cap, 2), dtype=torch.float64, device=dev)
        self.coord_3d = torch.empty((S, cap, 3), dtype=torch.float64, device=dev)
        self.conf = torch.empty((S, cap), dtype=torch.float32, device=dev)
        self.conf_obj = torch.empty((S, cap), dtype=torch.float32, device=dev)
        self.conf_frag = torch.empty((S, cap), dtype=torch.float32, device=dev)
        self.px = torch.empty((S, cap), dtype=torch.int32, device=dev)
        self.frag = torch.empty((S, cap), dtype=torch.int32, device=dev)
        self.counts = torch.zeros((S,), dtype=torch.int32, device=dev)
        self.totals = torch.zeros((S,), dtype=torch.int32, device=dev)
        self.offsets = (torch.arange(S, dtype=torch.int32, device=dev) * cap).contiguous()
        nbytes = _lib.lib().epos_corresp_workspace_bytes(B, J, h, w)
        self.workspace = torch.empty((nbytes,), dtype=torch.uint8, device=dev)


class CorrespExtractor:
    def __init__(self, device, num_objs, num_frags, model_store, obj_ids=None, output_scale=0.25, min_obj_conf=0.1,
                 min_frag_rel_conf=0.5, cap=4096, max_correspondences=4096):
        self.lib = _lib.lib()
        self.dev = torch.device(device)
        self.O, self.F = num_objs, num_frags
        self.obj_ids_list = list(model_store.dp_model['obj_ids']) if obj_ids is None else list(obj_ids)
        self.obj_ids = torch.tensor(self.obj_ids_list, dtype=torch.int32, device=self.dev)
        c, s = model_store.packed(num_objs)
        self.centers = torch.from_numpy(np.ascontiguousarray(c)).to(self.dev)
        self.sizes = torch.from_numpy(np.ascontiguousarray(s)).to(self.dev)
        self.output_scale, self.min_obj_conf, self.min_frag_rel_conf = output_scale, min_obj_conf, min_frag_rel_conf
        self.cap = cap
        self.max_corr = 0 if max_correspondences is None else int(max_correspondences)
        self._out = None

    def __call__(self, obj_conf, frag_conf, frag_loc, out=None, lazy_loc=None):
        """obj_conf [B,h,w,O+1], frag_conf [B,h,w,O,F], frag_loc [B,h,w,O,F,3] f32 CUDA -> BatchCorresp.
        frag_loc = None with lazy_loc = (feat_split [2,B*h*w,ld] bf16, w_loc [O*F*3,C] f32, b_loc [O*F*3] f32 or None):
        the localisation head is evaluated only at the surviving rows (epos_corresp_lazy_loc)."""
        B, h, w = obj_conf.shape[:3]
        J = len(self.obj_ids_list)
        if out is None:
            o = self._out
            if o is None or (o.B, o.J, o.cap, o.h, o.w) != (B, J, self.cap, h, w):
                o = self._out = BatchCorresp(self.dev, B, J, self.cap, h, w)
            out = o
        if frag_loc is None:
            feat, w_loc, b_loc = lazy_loc
            for t in (obj_conf, frag_conf, w_loc):
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
            assert feat.dtype == torch.bfloat16 and feat.dim() == 3 and feat.shape[1] == B * h * w and feat.stride(2) == 1
            _lib.check(self.lib.epos_corresp_lazy_loc(
                obj_conf.data_ptr(), frag_conf.data_ptr(), feat.data_ptr(), feat.stride(1), feat.stride(0), w_loc.shape[1],
                w_loc.data_ptr(), _lib.ptr(b_loc), B, h, w, self.O, self.F,
                self.obj_ids.data_ptr(), J, self.centers.data_ptr(), self.sizes.data_ptr(), float(self.output_scale),
                float(self.min_obj_conf), float(self.min_frag_rel_conf), self.cap, self.max_corr,
                out.coord_2d.data_ptr(), out.coord_3d.data_ptr(), out.conf.data_ptr(), out.conf_obj.data_ptr(),
                out.conf_frag.data_ptr(), out.px.data_ptr(), out.frag.data_ptr(), out.counts.data_ptr(),
                out.totals.data_ptr(), out.workspace.data_ptr(), out.workspace.numel(), _lib.stream_ptr()),
                'epos_corresp_lazy_loc')
            return out
        for t in (obj_conf, frag_conf, frag_loc):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        _lib.check(self.lib.epos_corresp(
            obj_conf.data_ptr(), frag_conf.data_ptr(), frag_loc.data_ptr(), B, h, w, self.O, self.F,
            self.obj_ids.data_ptr(), J, self.centers.data_ptr(), self.sizes.data_ptr(), float(self.output_scale),
            float(self.min_obj_conf), float(self.min_frag_rel_conf), self.cap, self.max_corr,
            out.coord_2d.data_ptr(), out.coord_3d.data_ptr(), out.conf.data_ptr(), out.conf_obj.data_ptr(),
            out.conf_frag.data_ptr(), out.px.data_ptr(), out.frag.data_ptr(), out.counts.data_ptr(),
            out.totals.data_ptr(), out.workspace.data_ptr(), out.workspace.numel(), _lib.stream_ptr()), 'epos_corresp')
        return out


def establish_many_to_many(obj_confs, frag_confs, frag_coords, gt_obj_ids, model_store, output_scale, min_obj_conf,
                           min_frag_rel_conf, project_to_surface=False, only_annotated_objs=False,
                           max_correspondences=None, cap=None):
    """Signature of corresp.establish_many_to_many (corresp.py:9-11) for ONE image: obj_confs [h,w,O+1],
    frag_confs [h,w,O,F], frag_coords [h,w,O,F,3] as torch CUDA tensors (or numpy arrays, copied to cuda:0).
    Returns {obj_id: {'px_id','frag_id','coord_2d','coord_3d','conf','conf_obj','conf_frag'}} of device tensors with
    the reference's row order.  'px_id' is the index into the object's masked-pixel list as in the reference."""
    if project_to_surface:
        raise NotImplementedError('project_to_surface needs igl and is off by default (infer.py:59-61)')
    dev = obj_confs.device if torch.is_tensor(obj_confs) else torch.device('cuda:0')

    def dv(a):
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return t.to(dev, torch.float32).contiguous()
    oc, fc, fl = dv(obj_confs)[None], dv(frag_confs)[None], dv(frag_coords)[None]
    h, w, O1 = oc.shape[1:]
    O, F = fc.shape[3], fc.shape[4]
    ids = [o for o in model_store.dp_model['obj_ids'] if not only_annotated_objs or o in gt_obj_ids]
    if not ids:
        return {}
    cap = cap or h * w * F
    ex = CorrespExtractor(dev, O, F, model_store, obj_ids=ids, output_scale=output_scale, min_obj_conf=min_obj_conf,
                          min_frag_rel_conf=min_frag_rel_conf, cap=min(cap, h * w * F),
                          max_correspondences=max_correspondences)
    bc = ex(oc, fc, fl)
    counts = bc.counts.cpu().numpy()
    out = {}
    for j, oid in enumerate(ids):
        n = int(counts[j])
        if n == 0:
            continue                                              # corresp.py:48-49
        px = bc.px[j, :n]
        mask = (oc[0, :, :, oid] > min_obj_conf).reshape(-1)
        rank = torch.cumsum(mask.to(torch.int64), 0) - 1          # position in the masked-pixel list
        out[oid] = {'px_id': rank[px.long()], 'frag_id': bc.frag[j, :n].long(), 'coord_2d': bc.coord_2d[j, :n],
                    'coord_3d': bc.coord_3d[j, :n], 'conf': bc.conf[j, :n], 'conf_obj': bc.conf_obj[j, :n],
                    'conf_frag': bc.conf_frag[j, :n], 'pixel': px.long()}
    return out
