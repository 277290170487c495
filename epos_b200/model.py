"""B200-native `model.predict` for EPOS (DeepLab-v3+ / Xception-65), drop-in for
/root/reference/epos_lib/model.py:629-687.

Host side only: PyTorch provides device memory and streams; every op is a hand-written sm_100a kernel
reached through the C ABI (include/epos_b200.h).  The layer sequence mirrors
  net_xception.py:396-483,593-657 (xception_65, output_stride 8 -> stride-to-atrous conversion :362-393),
  model.py:150-265 (ASPP), :268-393 (decoder), :396-458 (logit heads), :668-685 (softmax / argmax),
with inference BatchNorm folded into the conv weights/bias at load time and 1x1 convolutions executed
as split-bf16 (error-compensated) tcgen05 GEMMs.

Output contract (same keys / shapes / dtypes / channel order as the reference):
  pred_obj_conf  [B,h,w,O+1] f32 (softmax)      pred_obj_label [B,h,w] i64
  pred_frag_conf [B,h,w,O,F] f32 (softmax on F) pred_frag_loc  [B,h,w,O,F,3] f32 (raw)
as torch CUDA tensors.
"""
import collections

import numpy as np
import torch

from . import _lib
from .weights import RESNET50_BLOCKS, RN, XC, XCEPTION65_BLOCKS, head_channels

PRED_OBJ_CONF = 'pred_obj_conf'        # common.py:24-27
PRED_OBJ_LABEL = 'pred_obj_label'
PRED_FRAG_CONF = 'pred_frag_conf'
PRED_FRAG_LOC = 'pred_frag_loc'
LAZY_FRAG_LOC = 'lazy_frag_loc'        # engine path: (decoder features, f32 logit weights, bias) instead of pred_frag_loc

EPS_BACKBONE = 1e-3                    # feature.py:304
EPS_HEAD = 1e-5                        # model.py:197,310
EPS_RESNET = 1e-5                      # feature.py:277-281
RESNET_END_POINT = 'block1/unit_2/bottleneck_v1/conv3'   # feature.py:40-44
DECODER_END_POINT = 'entry_flow/block2/unit_1/xception_module/separable_conv2_pointwise'  # feature.py:61-66

_BLOCK_STRIDES = {'entry_flow/block1': 2, 'entry_flow/block2': 2, 'entry_flow/block3': 2,
                  'middle_flow/block1': 1, 'exit_flow/block1': 2, 'exit_flow/block2': 1}
_RELU_INSIDE = {'exit_flow/block2'}    # activation_fn_in_separable_conv=True, net_xception.py:640-647


class ModelOptions(collections.namedtuple('ModelOptions', [
        'outputs_to_num_channels', 'crop_size', 'atrous_rates', 'encoder_output_stride',
        'decoder_output_stride', 'model_variant', 'multi_grid'])):
    """Subset of common.ModelOptions (common.py:206-290) that the inference path reads."""
    __slots__ = ()

    def __new__(cls, outputs_to_num_channels, crop_size=(640, 480), atrous_rates=(12, 24, 36),
                encoder_output_stride=8, decoder_output_stride=(4,), model_variant='xception_65', multi_grid=None):
        return super().__new__(cls, outputs_to_num_channels, tuple(crop_size), tuple(atrous_rates),
                               encoder_output_stride, tuple(decoder_output_stride), model_variant,
                               tuple(multi_grid) if multi_grid else None)


def scale_dimension(dim, scale):
    """model.py:100-114."""
    return int((float(dim) - 1.0) * scale + 1.0)


def _bn_fold(w, scope, eps):
    g, b, m, v = (np.asarray(w['%s/BatchNorm/%s' % (scope, k)], np.float64)
                  for k in ('gamma', 'beta', 'moving_mean', 'moving_variance'))
    scale = g / np.sqrt(v + eps)
    return scale, b - m * scale


def _split(t, ld=None):
    """f32 [N,K] -> bf16 [2,N,ld] (hi, lo), zero-padded to a row pitch of ld >= K elements."""
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    out = torch.stack([hi, lo])
    if ld is not None and ld != t.shape[1]:
        out = torch.nn.functional.pad(out, (0, ld - t.shape[1]))
    return out.contiguous()


class Gemm:
    """Folded 1x1 conv: split-bf16 weights [2,N,K], f32 bias [N] (and f32 weights for the SIMT check path)."""

    def __init__(self, w_nk, bias, device, keep_f32):
        w32 = torch.from_numpy(np.ascontiguousarray(w_nk, dtype=np.float32)).to(device)
        self.N, self.K = w32.shape
        self.ldw = (self.K + 63) // 64 * 64          # weight rows padded to 128 B: aligned TMA rows (728 -> 768)
        self.w_split = _split(w32, self.ldw)
        self.w_f32 = w32 if keep_f32 else None
        self.bias = None if bias is None else torch.from_numpy(np.ascontiguousarray(bias, dtype=np.float32)).to(device)


class EposNet:
    def __init__(self, weights, num_objs, num_frags, device='cuda', model_options=None, keep_f32=False):
        self.lib = _lib.lib()
        self.dev = torch.device(device)
        self.O, self.F = num_objs, num_frags
        self.opts = model_options or ModelOptions(head_channels(num_objs, num_frags))
        if self.opts.model_variant not in ('xception_65', 'resnet_v1_50_beta') or self.opts.encoder_output_stride != 8:
            raise NotImplementedError('only xception_65 and resnet_v1_50_beta at output stride 8 are built')
        self.variant = self.opts.model_variant
        self.keep_f32 = keep_f32
        self.impl = 'tcgen05'            # 'simt' = fp32 validation path (needs keep_f32=True)
        self.end_points = {}
        self.gemm_events = None          # bench.py: list collecting (start, stop, M, N, K) per tcgen05 GEMM launch
        self._prepare(weights)

    # -- weight preparation ---------------------------------------------------------------------------
    def _dev(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.dev)

    def _conv_bn(self, w, scope, eps):
        k = np.asarray(w[scope + '/weights'], np.float64)                  # [kh,kw,Cin,Cout]
        scale, shift = _bn_fold(w, scope, eps)
        return k * scale[None, None, None, :], shift

    def _pw(self, w, scope, eps, keep_f32=None):
        k, shift = self._conv_bn(w, scope, eps)
        return Gemm(k[0, 0].T, shift, self.dev, self.keep_f32 if keep_f32 is None else keep_f32)   # [Cout, Cin]

    def _dw(self, w, scope, eps):
        k = np.asarray(w[scope + '/depthwise_weights'], np.float64)[:, :, :, 0]   # [3,3,C]
        scale, shift = _bn_fold(w, scope, eps)
        return self._dev((k * scale[None, None, :]).reshape(9, -1)), self._dev(shift)

    def _conv3(self, w, scope, eps):
        """3x3 conv + BN as an implicit-GEMM weight matrix [Cout, 9*Cin] with k = (ky*3+kx)*Cin + c."""
        k, shift = self._conv_bn(w, scope, eps)                                    # [3,3,Cin,Cout]
        return Gemm(k.reshape(-1, k.shape[3]).T, shift, self.dev, False)

    def _prepare_resnet(self, w, p):
        k, b = self._conv_bn(w, RN + '/conv1_1', EPS_RESNET)
        p['conv1_1'] = (self._dev(k), self._dev(b))
        p['conv1_2'] = self._conv3(w, RN + '/conv1_2', EPS_RESNET)
        p['conv1_3'] = self._conv3(w, RN + '/conv1_3', EPS_RESNET)
        cin = 128
        for scope, base_depth, units, _ in RESNET50_BLOCKS:
            for u in range(1, units + 1):
                base = '%s/%s/unit_%d/bottleneck_v1' % (RN, scope, u)
                if cin != base_depth * 4:
                    p[base + '/shortcut'] = self._pw(w, base + '/shortcut', EPS_RESNET)
                p[base + '/conv1'] = self._pw(w, base + '/conv1', EPS_RESNET)
                p[base + '/conv2'] = self._conv3(w, base + '/conv2', EPS_RESNET)
                p[base + '/conv3'] = self._pw(w, base + '/conv3', EPS_RESNET)
                cin = base_depth * 4

    def _prepare(self, w):
        p = {}
        if self.variant == 'resnet_v1_50_beta':
            self._prepare_resnet(w, p)
        else:
            self._prepare_xception(w, p)
        self._prepare_heads(w, p)
        self.p = p

    def _prepare_xception(self, w, p):
        # conv1_1 (3 -> 32) is emitted with 64 output channels, the upper 32 with zero weights and bias (ReLU(0) = 0), so
        # that conv1_2 (32 -> 64, zero-padded to Cin = 64) can run as an implicit 3x3 GEMM on the tensor cores.
        k, b = self._conv_bn(w, XC + '/entry_flow/conv1_1', EPS_BACKBONE)
        p['conv1_1'] = (self._dev(np.concatenate([k, np.zeros_like(k)], axis=3)),
                        self._dev(np.concatenate([b, np.zeros_like(b)])))
        k, b = self._conv_bn(w, XC + '/entry_flow/conv1_2', EPS_BACKBONE)
        k = np.concatenate([k, np.zeros_like(k)], axis=2)                           # [3,3,64,64]
        p['conv1_2'] = Gemm(k.reshape(-1, k.shape[3]).T, b, self.dev, False)
        for scope, depths, skip, units in XCEPTION65_BLOCKS:
            for u in range(1, units + 1):
                base = '%s/%s/unit_%d/xception_module' % (XC, scope, u)
                for i in range(3):
                    p['%s/dw%d' % (base, i)] = self._dw(w, '%s/separable_conv%d_depthwise' % (base, i + 1), EPS_BACKBONE)
                    p['%s/pw%d' % (base, i)] = self._pw(w, '%s/separable_conv%d_pointwise' % (base, i + 1), EPS_BACKBONE)
                if skip == 'conv':
                    p[base + '/shortcut'] = self._pw(w, base + '/shortcut', EPS_BACKBONE)

    def _prepare_heads(self, w, p):
        p['image_pooling'] = self._pw(w, 'image_pooling', EPS_HEAD, keep_f32=True)   # M = batch: fp32 SIMT kernel
        p['aspp0'] = self._pw(w, 'aspp0', EPS_HEAD)
        for i in (1, 2, 3):
            p['aspp%d_dw' % i] = self._dw(w, 'aspp%d_depthwise' % i, EPS_HEAD)
            p['aspp%d_pw' % i] = self._pw(w, 'aspp%d_pointwise' % i, EPS_HEAD)
        # concat_projection: input order [image_pooling, aspp0, aspp1, aspp2, aspp3] (model.py:233-256).
        k, shift = self._conv_bn(w, 'concat_projection', EPS_HEAD)
        wk = k[0, 0].T                                                      # [256, 1280]
        p['concat_proj_img'] = Gemm(wk[:, :256], shift, self.dev, True)     # image-level part -> per-image bias
        p['concat_proj'] = Gemm(wk[:, 256:], None, self.dev, self.keep_f32)
        p['feature_projection0'] = self._pw(w, 'decoder/feature_projection0', EPS_HEAD)
        for i in (0, 1):
            p['decoder_conv%d_dw' % i] = self._dw(w, 'decoder/decoder_conv%d_depthwise' % i, EPS_HEAD)
            p['decoder_conv%d_pw' % i] = self._pw(w, 'decoder/decoder_conv%d_pointwise' % i, EPS_HEAD)
        for name in (PRED_OBJ_CONF, PRED_FRAG_CONF, PRED_FRAG_LOC):
            k = np.asarray(w['logits/%s/weights' % name], np.float64)[0, 0].T
            # the localisation head keeps its f32 weights: the lazy head (engine path) evaluates it row by row in fp32
            p['logits/' + name] = Gemm(k, w['logits/%s/biases' % name], self.dev, self.keep_f32 or name == PRED_FRAG_LOC)

    # -- op wrappers ----------------------------------------------------------------------------------
    def _s(self):
        return torch.cuda.current_stream().cuda_stream

    def dwconv(self, x, B, H, W, C, ldx, wb, stride, rate, relu_in, relu_out, want_f32=False):
        Ho = H if stride == 1 else (H - 1) // 2 + 1
        Wo = W if stride == 1 else (W - 1) // 2 + 1
        ldy = (C + 15) // 16 * 16            # bf16 row pitch padded to 32 bytes (728 -> 736): sector-aligned rows
        y_split = torch.empty((2, B * Ho * Wo, ldy), dtype=torch.bfloat16, device=self.dev)
        y32 = torch.empty((B * Ho * Wo, C), dtype=torch.float32, device=self.dev) if want_f32 else None
        _lib.check(self.lib.epos_dwconv3x3(x.data_ptr(), ldx, wb[0].data_ptr(), wb[1].data_ptr(), _lib.ptr(y32),
                                           y_split.data_ptr(), ldy, B, H, W, C, stride, rate, int(relu_in), int(relu_out),
                                           self._s()), 'epos_dwconv3x3')
        return (y_split, y32, Ho, Wo) if want_f32 else (y_split, Ho, Wo)

    def split(self, x, B, H, W, C, ldx, subsample=1, relu=False):
        Ho, Wo = (H - 1) // subsample + 1, (W - 1) // subsample + 1
        ldy = (C + 15) // 16 * 16            # 32-byte aligned bf16 rows (see dwconv)
        y = torch.empty((2, B * Ho * Wo, ldy), dtype=torch.bfloat16, device=self.dev)
        _lib.check(self.lib.epos_split_bf16(x.data_ptr(), ldx, y.data_ptr(), ldy, y[0].numel(), B, H, W, C, subsample,
                                            int(relu), self._s()), 'epos_split_bf16')
        return y

    def gemm(self, a_split, g, M, relu, residual=None, out_f32=True, out_split=False, d_f32=None, ldd=None,
             d_split=None, ldd_split=None, bias=None, bias_group_rows=0, a_f32=None, pad_f32=True):
        """a_split [2,M,lda] bf16.  Returns (d_f32 or None, d_split or None)."""
        lda = a_split.shape[2]
        N, K = g.N, g.K
        bias_t = g.bias if bias is None else bias
        if out_f32 and d_f32 is None:
            # f32 activations with 128-byte aligned rows (728 -> 736 channels of pitch): the depthwise kernel's TMA boxes
            # (128 B of one pixel) then never straddle a line.  Only the 728-channel layers are affected.
            ldd = (N + 31) // 32 * 32 if (pad_f32 and N >= 64) else N
            d_f32 = torch.empty((M, ldd), dtype=torch.float32, device=self.dev)
        if out_split and d_split is None:
            d_split = torch.empty((2, M, N), dtype=torch.bfloat16, device=self.dev)
            ldd_split = N
        plane = 0 if d_split is None else d_split.stride(0)
        if self.impl == 'simt':
            # fp32 validation path: reconstruct A from the split planes, run the SIMT GEMM, then split.
            a32 = (a_split[0].float() + a_split[1].float()).contiguous()
            tmp = d_f32 if d_f32 is not None else torch.empty((M, N), dtype=torch.float32, device=self.dev)
            tl = ldd if d_f32 is not None else N
            _lib.check(self.lib.epos_pwconv_simt(a32.data_ptr(), a32.shape[1], g.w_f32.data_ptr(), _lib.ptr(bias_t),
                                                 bias_group_rows, _lib.ptr(residual),
                                                 0 if residual is None else residual.shape[-1], tmp.data_ptr(), tl,
                                                 M, N, K, int(relu), self._s()), 'epos_pwconv_simt')
            if d_split is not None:
                src = tmp if tl == N else torch.as_strided(tmp, (M, N), (tl, 1))
                hi = src.to(torch.bfloat16)
                torch.as_strided(d_split, (M, N), (ldd_split, 1), d_split.storage_offset()).copy_(hi)
                torch.as_strided(d_split, (M, N), (ldd_split, 1), d_split.storage_offset() + plane).copy_(
                    (src - hi.float()).to(torch.bfloat16))
            return d_f32, d_split
        ev = None
        if self.gemm_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _lib.check(self.lib.epos_pwconv_gemm(
            a_split.data_ptr(), lda, a_split.stride(0), g.w_split.data_ptr(), g.ldw, _lib.ptr(bias_t), bias_group_rows,
            _lib.ptr(residual), 0 if residual is None else residual.shape[-1],
            _lib.ptr(d_f32), ldd or 0, _lib.ptr(d_split), ldd_split or 0, plane, M, N, K, int(relu), self._s()),
            'epos_pwconv_gemm')
        if ev is not None:
            ev[1].record()
            self.gemm_events.append((ev[0], ev[1], M, N, K))
        return d_f32, d_split

    def conv3x3(self, x_split, g, B, H, W, C, rate, relu=True, out_f32=False, out_split=True):
        """x_split [2,B*H*W,C] bf16 -> 3x3 atrous conv + BN (+ReLU) as an implicit tcgen05 GEMM."""
        M, N = B * H * W, g.N
        d = torch.empty((M, N), dtype=torch.float32, device=self.dev) if out_f32 else None
        ds = torch.empty((2, M, N), dtype=torch.bfloat16, device=self.dev) if out_split else None
        ev = None
        if self.gemm_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _lib.check(self.lib.epos_conv3x3_gemm(
            x_split.data_ptr(), x_split.shape[2], x_split.stride(0), g.w_split.data_ptr(), g.ldw, _lib.ptr(g.bias), None, 0,
            _lib.ptr(d), N, _lib.ptr(ds), N, 0 if ds is None else ds.stride(0), B, H, W, C, N, rate, int(relu),
            self._s()), 'epos_conv3x3_gemm')
        if ev is not None:
            ev[1].record()
            self.gemm_events.append((ev[0], ev[1], M, N, 9 * C))
        return d, ds

    def bottleneck(self, x32, xs, B, H, W, cin, base, depth, depth_bottleneck, stride, rate, want_conv3=False):
        """net_resnet_v1_beta.py:38-93.  x32 f32 [M,cin] (None when only the split form exists), xs split-bf16.
        Returns (out f32, out split, Ho, Wo)."""
        p = self.p
        M = B * H * W
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        Mo = B * Ho * Wo
        if depth == cin:
            if stride == 1:
                shortcut = x32
            else:                                                 # resnet_utils.subsample
                shortcut = torch.empty((Mo, cin), dtype=torch.float32, device=self.dev)
                _lib.check(self.lib.epos_subsample_f32(x32.data_ptr(), cin, shortcut.data_ptr(), B, H, W, cin, stride,
                                                       self._s()), 'epos_subsample_f32')
        else:
            xsc = xs if stride == 1 else self.split(x32, B, H, W, cin, x32.shape[1], subsample=stride)
            shortcut, _ = self.gemm(xsc, p[base + '/shortcut'], Mo, relu=False)
        _, r1 = self.gemm(xs, p[base + '/conv1'], M, relu=True, out_f32=False, out_split=True)
        if stride == 1:
            _, r2 = self.conv3x3(r1, p[base + '/conv2'], B, H, W, depth_bottleneck, rate)
        else:
            # conv2d_same(stride 2) == the stride-1 SAME conv sampled at even pixels (resnet_v1_test.py:72-149)
            full, _ = self.conv3x3(r1, p[base + '/conv2'], B, H, W, depth_bottleneck, rate, out_f32=True, out_split=False)
            r2 = self.split(full, B, H, W, depth_bottleneck, depth_bottleneck, subsample=stride)
        if want_conv3:                                            # end point = conv3 + BN before the residual add
            c3, _ = self.gemm(r2, p[base + '/conv3'], Mo, relu=False)
            self.end_points[base.replace(RN + '/', '') + '/conv3'] = (c3, Ho, Wo, depth)
        out, outs = self.gemm(r2, p[base + '/conv3'], Mo, relu=True, residual=shortcut, out_f32=True, out_split=True)
        return out, outs, Ho, Wo

    def resnet_features(self, images):
        """resnet_v1_50_beta at output stride 8 (net_resnet_v1_beta.py:302-373; resnet_utils.py:125-217)."""
        lib, p = self.lib, self.p
        B, H, W, _ = images.shape
        H1, W1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        c1 = torch.empty((2, B * H1 * W1, 64), dtype=torch.bfloat16, device=self.dev)
        _lib.check(lib.epos_conv3x3_rgb_s2(images.data_ptr(), p['conv1_1'][0].data_ptr(), p['conv1_1'][1].data_ptr(),
                                           None, c1.data_ptr(), B, H, W, 64, self._s()), 'epos_conv3x3_rgb_s2')
        _, c2 = self.conv3x3(c1, p['conv1_2'], B, H1, W1, 64, 1)
        c3, _ = self.conv3x3(c2, p['conv1_3'], B, H1, W1, 64, 1, out_f32=True, out_split=False)
        h, w = (H1 + 1) // 2, (W1 + 1) // 2
        x32 = torch.empty((B * h * w, 128), dtype=torch.float32, device=self.dev)
        xs = torch.empty((2, B * h * w, 128), dtype=torch.bfloat16, device=self.dev)
        _lib.check(lib.epos_maxpool3x3_s2(c3.data_ptr(), x32.data_ptr(), xs.data_ptr(), B, H1, W1, 128, self._s()),
                   'epos_maxpool3x3_s2')
        cin = 128
        target = self.opts.encoder_output_stride // 4              # net_resnet_v1_beta.py:183-185
        current_stride, rate = 1, 1
        mg = self.opts.multi_grid or (1, 1, 1)
        for scope, base_depth, units, last_stride in RESNET50_BLOCKS:
            for u in range(1, units + 1):
                base = '%s/%s/unit_%d/bottleneck_v1' % (RN, scope, u)
                stride = last_stride if u == units else 1
                unit_rate = mg[u - 1] if scope == 'block4' else 1
                want = base.endswith(RESNET_END_POINT.rsplit('/', 1)[0])
                if current_stride == target:                       # resnet_utils.py:191-197
                    x32, xs, h, w = self.bottleneck(x32, xs, B, h, w, cin, base, base_depth * 4, base_depth, 1,
                                                    rate * unit_rate, want)
                    rate *= stride
                else:
                    x32, xs, h, w = self.bottleneck(x32, xs, B, h, w, cin, base, base_depth * 4, base_depth, stride,
                                                    unit_rate, want)
                    current_stride *= stride
                cin = base_depth * 4
        return x32, xs, h, w, cin

    def small_fc(self, a, w_f32, bias, relu):
        M, K = a.shape
        N = w_f32.shape[0]
        d = torch.empty((M, N), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.epos_pwconv_simt(a.data_ptr(), K, w_f32.data_ptr(), _lib.ptr(bias), 0, None, 0,
                                             d.data_ptr(), N, M, N, K, int(relu), self._s()), 'epos_pwconv_simt')
        return d

    # -- network --------------------------------------------------------------------------------------
    def xception_module(self, x, B, H, W, cin, base, depths, skip, stride, rate, relu_inside):
        """x: f32 [B*H*W, cin].  net_xception.py:198-323."""
        r, c, h, w = x, cin, H, W
        for i in range(3):
            s = stride if i == 2 else 1
            a, ho, wo = self.dwconv(r, B, h, w, c, r.shape[1], self.p['%s/dw%d' % (base, i)], s, rate,
                                    relu_in=not relu_inside, relu_out=relu_inside)
            M = B * ho * wo
            g = self.p['%s/pw%d' % (base, i)]
            res = None
            if i == 2 and skip == 'conv':
                xs = self.split(x, B, H, W, cin, x.shape[1], subsample=stride)
                res, _ = self.gemm(xs, self.p[base + '/shortcut'], M, relu=False)
            elif i == 2 and skip == 'sum':
                res = x
            r, _ = self.gemm(a, g, M, relu=relu_inside, residual=res)
            c, h, w = depths[i], ho, wo
            self.end_points['%s/separable_conv%d_pointwise' % (base.replace(XC + '/', ''), i + 1)] = (r, h, w, c)
        return r, h, w, c

    def forward_features(self, images):
        """images [B,H,W,3] f32 cuda in [0,255] -> decoder features (split-bf16 [2,M,256]) and (B,h,w)."""
        lib, p = self.lib, self.p
        B, H, W, _ = images.shape
        images = images.contiguous()
        self.end_points = {}
        if self.variant == 'resnet_v1_50_beta':
            x, xs, h, w, c = self.resnet_features(images)
            return self.aspp_decoder(x, xs, B, H, W, h, w, c, RESNET_END_POINT)
        H1, W1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        c1 = torch.empty((2, B * H1 * W1, 64), dtype=torch.bfloat16, device=self.dev)
        _lib.check(lib.epos_conv3x3_rgb_s2(images.data_ptr(), p['conv1_1'][0].data_ptr(), p['conv1_1'][1].data_ptr(),
                                           None, c1.data_ptr(), B, H, W, 64, self._s()), 'epos_conv3x3_rgb_s2')
        c2, _ = self.conv3x3(c1, p['conv1_2'], B, H1, W1, 64, 1, relu=True, out_f32=True, out_split=False)
        x, h, w, c = c2, H1, W1, 64
        target = self.opts.encoder_output_stride // 2              # net_xception.py:455-458
        current_stride, rate = 1, 1
        for scope, depths, skip, units in XCEPTION65_BLOCKS:
            stride = _BLOCK_STRIDES[scope]
            for u in range(1, units + 1):
                base = '%s/%s/unit_%d/xception_module' % (XC, scope, u)
                if current_stride == target:                       # net_xception.py:374-385
                    x, h, w, c = self.xception_module(x, B, h, w, c, base, depths, skip, 1, rate, scope in _RELU_INSIDE)
                    rate *= stride
                else:
                    x, h, w, c = self.xception_module(x, B, h, w, c, base, depths, skip, stride, 1, scope in _RELU_INSIDE)
                    current_stride *= stride
        return self.aspp_decoder(x, None, B, H, W, h, w, c, DECODER_END_POINT)

    def aspp_decoder(self, x, xs, B, H, W, h, w, c, end_point):
        """ASPP + decoder on backbone features x (f32 [B*h*w, c]; xs = its split form if already available)."""
        lib, p = self.lib, self.p
        self.end_points['backbone'] = (x, h, w, c)
        # ---- ASPP (model.py:217-258) ----
        M = B * h * w
        pooled = torch.empty((B, c), dtype=torch.float32, device=self.dev)
        _lib.check(lib.epos_global_mean(x.data_ptr(), pooled.data_ptr(), B, h * w, c, self._s()), 'epos_global_mean')
        ip = self.small_fc(pooled, p['image_pooling'].w_f32, p['image_pooling'].bias, relu=True)       # [B,256]
        cp_bias = self.small_fc(ip, p['concat_proj_img'].w_f32, p['concat_proj_img'].bias, relu=False)  # [B,256]
        nb = 1 + len(self.opts.atrous_rates)
        cat = torch.empty((2, M, 256 * nb), dtype=torch.bfloat16, device=self.dev)
        if xs is None:
            xs = self.split(x, B, h, w, c, c)
        self.gemm(xs, p['aspp0'], M, relu=True, out_f32=False, d_split=cat, ldd_split=256 * nb)
        for i, r_ in enumerate(self.opts.atrous_rates, 1):
            a, _, _ = self.dwconv(x, B, h, w, c, x.shape[1], p['aspp%d_dw' % i], 1, r_, relu_in=False, relu_out=True)
            self.gemm(a, p['aspp%d_pw' % i], M, relu=True, out_f32=False, d_split=cat[:, :, 256 * i:],
                      ldd_split=256 * nb)
        aspp, _ = self.gemm(cat, p['concat_proj'], M, relu=True, bias=cp_bias, bias_group_rows=h * w)
        self.end_points['aspp'] = (aspp, h, w, 256)
        # ---- decoder (model.py:325-380) ----
        skip, sh, sw, sc = self.end_points[end_point]
        dstride = self.opts.decoder_output_stride[0]
        dw_ = scale_dimension(W, 1.0 / dstride)     # crop_size == image size at inference (infer.py:650-654)
        dh_ = scale_dimension(H, 1.0 / dstride)
        if (sh, sw) != (dh_, dw_):
            raise NotImplementedError('skip feature %dx%d != decoder size %dx%d' % (sh, sw, dh_, dw_))
        Md = B * dh_ * dw_
        dcat = torch.empty((Md, 320), dtype=torch.float32, device=self.dev)       # 304 channels, 128-byte aligned rows
        ss = self.split(skip, B, sh, sw, sc, skip.shape[1])
        self.gemm(ss, p['feature_projection0'], Md, relu=True, d_f32=dcat[:, 256:], ldd=320)
        _lib.check(lib.epos_resize_bilinear(aspp.data_ptr(), dcat.data_ptr(), 320, B, h, w, dh_, dw_, 256, self._s()),
                   'epos_resize_bilinear')
        a, _, _ = self.dwconv(dcat, B, dh_, dw_, 304, 320, p['decoder_conv0_dw'], 1, 1, False, True)
        y0, _ = self.gemm(a, p['decoder_conv0_pw'], Md, relu=True)
        a, _, _ = self.dwconv(y0, B, dh_, dw_, 256, 256, p['decoder_conv1_dw'], 1, 1, False, True)
        y1, y1s = self.gemm(a, p['decoder_conv1_pw'], Md, relu=True, out_f32=self.keep_f32, out_split=True)
        self.end_points['decoder'] = (y1, dh_, dw_, 256)
        return y1s, B, dh_, dw_

    def heads(self, feat_split, B, h, w, lazy_loc=False, sparse_conf=None):
        """Logit heads + softmax/argmax (model.py:448-456, 676-685).  Materialising mode (the drop-in contract of
        model.predict) writes all four maps.  lazy_loc=True (engine path) skips the pred_frag_loc GEMM -- 3 O F columns,
        2.4 GB per image at O = 30 / F = 256, read back at <= max_correspondences rows per object -- and returns the
        decoder features and the f32 logit weights instead; corresp.CorrespExtractor evaluates the head at the
        surviving rows only.  sparse_conf = min_obj_conf (engine path, F != 64): the fragment softmax runs only on the
        (pixel, object) pairs whose object confidence exceeds it -- the pairs establish_many_to_many reads; the other
        rows of pred_frag_conf keep their logits."""
        M = B * h * w
        O, F = self.O, self.F
        obj, _ = self.gemm(feat_split, self.p['logits/' + PRED_OBJ_CONF], M, relu=False, pad_f32=False)
        # softmax over F in the GEMM epilogue (relu = 2) when a fragment group is exactly one 64-column group;
        # other F (e.g. config 5's 256) use the row-softmax kernel
        fused = F == 64 and self.impl == 'tcgen05'
        fc, _ = self.gemm(feat_split, self.p['logits/' + PRED_FRAG_CONF], M, relu=2 if fused else False, pad_f32=False)
        labels = torch.empty((M,), dtype=torch.int64, device=self.dev)
        _lib.check(self.lib.epos_softmax_rows(obj.data_ptr(), labels.data_ptr(), M, O + 1, self._s()), 'epos_softmax_rows')
        if not fused and sparse_conf is not None:
            _lib.check(self.lib.epos_softmax_rows_masked(fc.data_ptr(), obj.data_ptr(), M, O, F, float(sparse_conf), self._s()),
                       'epos_softmax_rows_masked')
        elif not fused:
            _lib.check(self.lib.epos_softmax_rows(fc.data_ptr(), None, M * O, F, self._s()), 'epos_softmax_rows')
        out = {PRED_OBJ_CONF: obj.view(B, h, w, O + 1), PRED_OBJ_LABEL: labels.view(B, h, w),
               PRED_FRAG_CONF: fc.view(B, h, w, O, F)}
        g = self.p['logits/' + PRED_FRAG_LOC]
        if lazy_loc:
            out[LAZY_FRAG_LOC] = (feat_split, g.w_f32, g.bias)
        else:
            fl, _ = self.gemm(feat_split, g, M, relu=False, pad_f32=False)
            out[PRED_FRAG_LOC] = fl.view(B, h, w, O, F, 3)
        return out

    def predict(self, images, lazy_loc=False, sparse_conf=None):
        feat, B, h, w = self.forward_features(images)
        return self.heads(feat, B, h, w, lazy_loc, sparse_conf)


_NETS = {}
_NETS_MAX = 2


def predict(images, model_options=None, upsample_logits=False, image_pyramid=None, num_objs=None, num_frags=None,
            frag_cls_agnostic=False, frag_loc_agnostic=False, weights=None, net=None):
    """Signature of model.predict (model.py:629-687) + `weights`/`net` (the reference reads variables from
    the TF graph; here they are passed explicitly).  Returns a dict of torch CUDA tensors."""
    if upsample_logits or image_pyramid or frag_cls_agnostic or frag_loc_agnostic:
        raise NotImplementedError('only the default inference configuration of scripts/infer.py is built')
    if net is None:
        # the cache entry holds a reference to the weights dict, so its id cannot be recycled while the entry lives;
        # one network per (weights, head shape, options, device), at most _NETS_MAX kept (oldest evicted)
        key = (id(weights), num_objs, num_frags, model_options, str(images.device))
        hit = _NETS.get(key)
        if hit is None or hit[0] is not weights:
            hit = (weights, EposNet(weights, num_objs, num_frags, images.device, model_options))
            _NETS[key] = hit
            while len(_NETS) > _NETS_MAX:
                _NETS.pop(next(iter(_NETS)))
        net = hit[1]
    return net.predict(images)
