"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the EPOS
DeepLab-v3+/Xception-65 forward pass `model.predict`.

Follows, line by line in behaviour (not in code):
  /root/reference/epos_lib/model.py:629-687      predict (softmax / argmax / reshape)
  /root/reference/epos_lib/model.py:150-265      extract_encoder_features (ASPP)
  /root/reference/epos_lib/model.py:268-393      extract_decoder_features
  /root/reference/epos_lib/model.py:396-458      get_branch_logits
  /root/reference/epos_lib/feature.py:171-174    _preprocess_zero_mean_unit_range
  /root/reference/epos_lib/net_xception.py:74-194   fixed_padding / separable_conv2d_same
  /root/reference/epos_lib/net_xception.py:198-323  xception_module
  /root/reference/epos_lib/net_xception.py:327-393  stack_blocks_dense (stride -> atrous rate)
  /root/reference/epos_lib/net_xception.py:396-483,593-657  xception / xception_65
  /root/reference/external/slim/nets/resnet_utils.py:77-122  conv2d_same
  /root/reference/epos_lib/misc.py:94-107        resize_bilinear(align_corners=True)
  /root/reference/epos_lib/net_resnet_v1_beta.py:38-93,96-112,168-204,302-373  bottleneck, beta root, resnet_v1_50_beta
  /root/reference/external/slim/nets/resnet_utils.py:59-74,125-217  subsample, stack_blocks_dense

Third-party arithmetic that is not under /root/reference (TensorFlow 1.12 kernels) is restated
with PyTorch CPU ops: conv (TF 'SAME' = total pad max((ceil(n/s)-1)*s+k_eff-n,0), before =
floor(total/2)), inference BatchNorm gamma*(x-mu)/sqrt(var+eps)+beta, bilinear resize with
align_corners.  Pinned by the Slim golden vectors (tests/test_oracle_cnn.py); the network as a
whole has no reference test => "parity unpinned" beyond those vectors.

Weights are a dict of numpy arrays keyed by TF variable names, TF layouts (HWIO, [kh,kw,C,1]).
"""
import numpy as np
import torch
import torch.nn.functional as F

XC = 'xception_65'
EPS_BACKBONE = 1e-3   # feature.py:304 (xception arg scope)
EPS_HEAD = 1e-5       # model.py:197,310

BLOCKS = [  # scope, depth_list, skip, units, stride, relu-inside-separable-conv
    ('entry_flow/block1', [128, 128, 128], 'conv', 1, 2, False),
    ('entry_flow/block2', [256, 256, 256], 'conv', 1, 2, False),
    ('entry_flow/block3', [728, 728, 728], 'conv', 1, 2, False),
    ('middle_flow/block1', [728, 728, 728], 'sum', 16, 1, False),
    ('exit_flow/block1', [728, 1024, 1024], 'conv', 1, 2, False),
    ('exit_flow/block2', [1536, 1536, 2048], 'none', 1, 1, True),
]
DECODER_END_POINT = 'entry_flow/block2/unit_1/xception_module/separable_conv2_pointwise'

RN = 'resnet_v1_50'   # name_scope['resnet_v1_50_beta'], feature.py:140-150
EPS_RESNET = 1e-5     # feature.py:277-281
RESNET50_BLOCKS = [('block1', 64, 3, 2), ('block2', 128, 4, 2), ('block3', 256, 6, 2), ('block4', 512, 3, 1)]
RESNET_END_POINT = 'block1/unit_2/bottleneck_v1/conv3'   # feature.py:40-44


def scale_dimension(dim, scale):
    """model.py:100-114."""
    return int((float(dim) - 1.0) * scale + 1.0)


def tf_same_pad(n, k_eff, s):
    total = max((-(-n // s) - 1) * s + k_eff - n, 0)
    return total // 2, total - total // 2


class Oracle:
    def __init__(self, weights, dtype=torch.float32, threads=None, model_variant='xception_65', multi_grid=None):
        self.w = weights
        self.dt = dtype
        self.variant = model_variant
        self.multi_grid = tuple(multi_grid) if multi_grid else (1, 1, 1)   # net_resnet_v1_beta.py:36
        if threads:
            torch.set_num_threads(threads)
        self.end_points = {}

    # -- primitive restatements -------------------------------------------------------------
    def t(self, name):
        return torch.from_numpy(np.ascontiguousarray(self.w[name])).to(self.dt)

    def bn(self, x, scope, eps):
        g, b, m, v = (self.t('%s/BatchNorm/%s' % (scope, k)).view(1, -1, 1, 1)
                      for k in ('gamma', 'beta', 'moving_mean', 'moving_variance'))
        return g * (x - m) / torch.sqrt(v + eps) + b

    def conv(self, x, scope, stride=1, rate=1, padding='SAME'):
        """slim.conv2d without normaliser; TF padding semantics; weights HWIO."""
        w = self.t(scope + '/weights').permute(3, 2, 0, 1).contiguous()
        k = w.shape[-1]
        if padding == 'SAME':
            k_eff = k + (k - 1) * (rate - 1)
            ph = tf_same_pad(x.shape[2], k_eff, stride)
            pw = tf_same_pad(x.shape[3], k_eff, stride)
            x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]))
        return F.conv2d(x, w, stride=stride, dilation=rate)

    def depthwise(self, x, scope, stride=1, rate=1, padding='SAME'):
        w = self.t(scope + '/depthwise_weights')            # [kh,kw,C,1]
        c = w.shape[2]
        w = w.permute(2, 3, 0, 1).contiguous()              # [C,1,kh,kw]
        k = w.shape[-1]
        if padding == 'SAME':
            k_eff = k + (k - 1) * (rate - 1)
            ph = tf_same_pad(x.shape[2], k_eff, stride)
            pw = tf_same_pad(x.shape[3], k_eff, stride)
            x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]))
        return F.conv2d(x, w, stride=stride, dilation=rate, groups=c)

    @staticmethod
    def fixed_padding(x, k, rate):
        """net_xception.py:74-93."""
        k_eff = k + (k - 1) * (rate - 1)
        tot = k_eff - 1
        beg = tot // 2
        return F.pad(x, (beg, tot - beg, beg, tot - beg))

    def conv2d_same(self, x, scope, stride, rate=1):
        """resnet_utils.conv2d_same + BN + ReLU (xception arg scope)."""
        if stride == 1:
            y = self.conv(x, scope, 1, rate, 'SAME')
        else:
            y = self.conv(self.fixed_padding(x, 3, rate), scope, stride, rate, 'VALID')
        return F.relu(self.bn(y, scope, EPS_BACKBONE))

    def separable_same(self, x, scope, stride, rate, relu_after):
        """separable_conv2d_same, split form (regularize_depthwise=False): dw,BN[,ReLU],pw,BN[,ReLU]."""
        if stride == 1:
            y = self.depthwise(x, scope + '_depthwise', 1, rate, 'SAME')
        else:
            y = self.depthwise(self.fixed_padding(x, 3, rate), scope + '_depthwise', stride, rate, 'VALID')
        y = self.bn(y, scope + '_depthwise', EPS_BACKBONE)
        if relu_after:
            y = F.relu(y)
        y = self.conv(y, scope + '_pointwise')
        y = self.bn(y, scope + '_pointwise', EPS_BACKBONE)
        if relu_after:
            y = F.relu(y)
        self.end_points[scope.replace(XC + '/', '') + '_pointwise'] = y
        return y

    def xception_module(self, x, base, depths, skip, stride, rate, relu_inside):
        r = x
        for i in range(3):
            if not relu_inside:
                r = F.relu(r)                                   # net_xception.py:276
            r = self.separable_same(r, '%s/separable_conv%d' % (base, i + 1),
                                    stride if i == 2 else 1, rate, relu_inside)
        if skip == 'conv':
            s = self.conv(x, base + '/shortcut', stride, 1, 'SAME')
            s = self.bn(s, base + '/shortcut', EPS_BACKBONE)
            return r + s
        if skip == 'sum':
            return r + x
        return r

    def split_separable(self, x, scope, rate):
        """model.py:51-97 (dw,BN,ReLU,pw,BN,ReLU; eps 1e-5)."""
        y = self.depthwise(x, scope + '_depthwise', 1, rate, 'SAME')
        y = F.relu(self.bn(y, scope + '_depthwise', EPS_HEAD))
        y = self.conv(y, scope + '_pointwise')
        return F.relu(self.bn(y, scope + '_pointwise', EPS_HEAD))

    def conv_bn_relu_head(self, x, scope):
        return F.relu(self.bn(self.conv(x, scope), scope, EPS_HEAD))

    # -- ResNet-50 beta -------------------------------------------------------------------
    def rn_conv(self, x, scope, k, stride=1, rate=1, relu=True):
        """slim.conv2d / resnet_utils.conv2d_same under resnet_arg_scope: conv + BN(eps 1e-5) (+ReLU)."""
        if k == 1:
            y = self.conv(x, scope, stride, 1, 'SAME')              # 1x1: SAME at stride s samples every s-th pixel
        elif stride == 1:
            y = self.conv(x, scope, 1, rate, 'SAME')
        else:
            y = self.conv(self.fixed_padding(x, k, rate), scope, stride, rate, 'VALID')
        y = self.bn(y, scope, EPS_RESNET)
        return F.relu(y) if relu else y

    def bottleneck(self, x, base, depth, depth_bottleneck, stride, rate):
        """net_resnet_v1_beta.py:38-93."""
        if depth == x.shape[1]:
            shortcut = x if stride == 1 else x[:, :, ::stride, ::stride]      # subsample = 1x1 max-pool, stride s
        else:
            shortcut = self.rn_conv(x, base + '/shortcut', 1, stride, relu=False)
        r = self.rn_conv(x, base + '/conv1', 1)
        r = self.rn_conv(r, base + '/conv2', 3, stride, rate)
        r = self.rn_conv(r, base + '/conv3', 1, relu=False)
        self.end_points[base.replace(RN + '/', '') + '/conv3'] = r
        return F.relu(shortcut + r)

    def resnet_backbone(self, images_nhwc, output_stride=8):
        x = torch.from_numpy(np.ascontiguousarray(images_nhwc)).to(self.dt).permute(0, 3, 1, 2)
        x = (2.0 / 255.0) * x - 1.0                             # beta variants: feature.py:176-185
        x = self.rn_conv(x, RN + '/conv1_1', 3, 2)              # root_block_fn_for_beta_variant
        x = self.rn_conv(x, RN + '/conv1_2', 3, 1)
        x = self.rn_conv(x, RN + '/conv1_3', 3, 1)
        ph = tf_same_pad(x.shape[2], 3, 2)                      # slim.max_pool2d(3, 2, 'SAME')
        pw = tf_same_pad(x.shape[3], 3, 2)
        x = F.max_pool2d(F.pad(x, (pw[0], pw[1], ph[0], ph[1]), value=float('-inf')), 3, 2)
        target = output_stride // 4                             # net_resnet_v1_beta.py:183-185
        current_stride, rate = 1, 1
        for scope, base_depth, units, last_stride in RESNET50_BLOCKS:
            for u in range(1, units + 1):
                base = '%s/%s/unit_%d/bottleneck_v1' % (RN, scope, u)
                stride = last_stride if u == units else 1
                unit_rate = self.multi_grid[u - 1] if scope == 'block4' else 1
                if current_stride == target:                    # resnet_utils.py:191-197
                    x = self.bottleneck(x, base, base_depth * 4, base_depth, 1, rate * unit_rate)
                    rate *= stride
                else:
                    x = self.bottleneck(x, base, base_depth * 4, base_depth, stride, unit_rate)
                    current_stride *= stride
        return x

    # -- network ----------------------------------------------------------------------------
    def backbone(self, images_nhwc, output_stride=8):
        if self.variant == 'resnet_v1_50_beta':
            return self.resnet_backbone(images_nhwc, output_stride)
        x = torch.from_numpy(np.ascontiguousarray(images_nhwc)).to(self.dt).permute(0, 3, 1, 2)
        x = (2.0 / 255.0) * x - 1.0                             # feature.py:171-174
        x = self.conv2d_same(x, XC + '/entry_flow/conv1_1', 2)
        x = self.conv2d_same(x, XC + '/entry_flow/conv1_2', 1)
        target = output_stride // 2                             # net_xception.py:455-458
        current_stride, rate = 1, 1
        for scope, depths, skip, units, stride, relu_inside in BLOCKS:
            for u in range(1, units + 1):
                base = '%s/%s/unit_%d/xception_module' % (XC, scope, u)
                if current_stride == target:                    # net_xception.py:374-385
                    x = self.xception_module(x, base, depths, skip, 1, rate, relu_inside)
                    rate *= stride
                else:
                    x = self.xception_module(x, base, depths, skip, stride, 1, relu_inside)
                    current_stride *= stride
        return x

    def aspp(self, f, atrous_rates=(12, 24, 36)):
        h, w = f.shape[2:]
        img = f.mean(dim=(2, 3), keepdim=True)                  # model.py:220
        img = self.conv_bn_relu_head(img, 'image_pooling')
        img = F.interpolate(img, size=(h, w), mode='bilinear', align_corners=True)
        branches = [img, self.conv_bn_relu_head(f, 'aspp0')]
        for i, r in enumerate(atrous_rates, 1):
            branches.append(self.split_separable(f, 'aspp%d' % i, r))
        x = torch.cat(branches, dim=1)
        return self.conv_bn_relu_head(x, 'concat_projection')  # dropout inactive

    def decoder(self, x, crop_size_wh, decoder_stride=4):
        skip = self.end_points[RESNET_END_POINT if self.variant == 'resnet_v1_50_beta' else DECODER_END_POINT]
        proj = self.conv_bn_relu_head(skip, 'decoder/feature_projection0')
        dw_ = scale_dimension(crop_size_wh[0], 1.0 / decoder_stride)
        dh_ = scale_dimension(crop_size_wh[1], 1.0 / decoder_stride)
        feats = [F.interpolate(t_, size=(dh_, dw_), mode='bilinear', align_corners=True)
                 for t_ in (x, proj)]
        y = torch.cat(feats, dim=1)
        y = self.split_separable(y, 'decoder/decoder_conv0', 1)
        return self.split_separable(y, 'decoder/decoder_conv1', 1)

    def logits(self, x, name):
        w = self.t('logits/%s/weights' % name).permute(3, 2, 0, 1).contiguous()
        return F.conv2d(x, w, bias=self.t('logits/%s/biases' % name))

    def predict(self, images_nhwc, num_objs, num_frags, return_features=False):
        """Returns the dict of model.predict with numpy NHWC arrays (batch dim kept)."""
        with torch.no_grad():
            self.end_points = {}
            h, w = images_nhwc.shape[1:3]
            f = self.backbone(images_nhwc)
            a = self.aspp(f)
            d = self.decoder(a, (w, h))
            out = {}
            b, _, oh, ow = d.shape
            lo = self.logits(d, 'pred_obj_conf').permute(0, 2, 3, 1)
            lc = self.logits(d, 'pred_frag_conf').permute(0, 2, 3, 1).reshape(b, oh, ow, num_objs, num_frags)
            ll = self.logits(d, 'pred_frag_loc').permute(0, 2, 3, 1).reshape(b, oh, ow, num_objs, num_frags, 3)
            obj_conf = torch.softmax(lo, dim=-1)
            out['pred_obj_conf'] = obj_conf.contiguous().numpy()
            out['pred_obj_label'] = torch.argmax(obj_conf, dim=-1).numpy().astype(np.int64)
            out['pred_frag_conf'] = torch.softmax(lc, dim=-1).contiguous().numpy()
            out['pred_frag_loc'] = ll.contiguous().numpy()
            if return_features:
                out['_backbone'] = f.permute(0, 2, 3, 1).contiguous().numpy()
                out['_aspp'] = a.permute(0, 2, 3, 1).contiguous().numpy()
                out['_decoder'] = d.permute(0, 2, 3, 1).contiguous().numpy()
                out['_obj_logits'] = lo.contiguous().numpy()
                ep = RESNET_END_POINT if self.variant == 'resnet_v1_50_beta' else DECODER_END_POINT
                out['_skip'] = self.end_points[ep].permute(0, 2, 3, 1).contiguous().numpy()
            return out


def predict(weights, images_nhwc, num_objs, num_frags, dtype=torch.float32, model_variant='xception_65',
            multi_grid=None, **kw):
    return Oracle(weights, dtype, model_variant=model_variant, multi_grid=multi_grid).predict(
        images_nhwc, num_objs, num_frags, **kw)
