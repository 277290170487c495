"""Weight container for the EPOS DeepLab-v3+/Xception-65 network, keyed by TF variable names.

The reference restores a TF-1.12 checkpoint whose variable names are `var.op.name`
(/root/reference/epos_lib/misc.py:159-168, scripts/infer.py:670-683).  This module keeps
exactly those names so that a checkpoint dumped to .npz is usable without a renaming table,
and provides the "random-init" generator that mirrors the reference initialisers:

  * backbone conv / depthwise / pointwise: truncated_normal(stddev=0.09)
        (/root/reference/epos_lib/net_xception.py:745,782-783)
  * ASPP / decoder split-separable: depthwise stddev 0.33, pointwise stddev 0.06
        (/root/reference/epos_lib/model.py:51-97)
  * ASPP 1x1, image pooling, concat projection, feature projection: slim default
        Xavier-uniform (slim.conv2d default `weights_initializer`)
  * logits: truncated_normal(stddev=0.01), zero bias (/root/reference/epos_lib/model.py:437)
  * BatchNorm at init: gamma=1, beta=0, moving_mean=0, moving_variance=1

Layouts are TF's: conv `HWIO` [kh,kw,Cin,Cout], depthwise [kh,kw,C,1].
"""
import numpy as np

XC = 'xception_65'
BN_KEYS = ('gamma', 'beta', 'moving_mean', 'moving_variance')

# (block scope, depth_list, skip type, num_units)   -- net_xception.py:604-647
XCEPTION65_BLOCKS = [
    ('entry_flow/block1', [128, 128, 128], 'conv', 1),
    ('entry_flow/block2', [256, 256, 256], 'conv', 1),
    ('entry_flow/block3', [728, 728, 728], 'conv', 1),
    ('middle_flow/block1', [728, 728, 728], 'sum', 16),
    ('exit_flow/block1', [728, 1024, 1024], 'conv', 1),
    ('exit_flow/block2', [1536, 1536, 2048], 'none', 1),
]


RN = 'resnet_v1_50'               # name_scope['resnet_v1_50_beta'] (feature.py:140-150)
# (block scope, base depth, num_units, stride of the last unit)   -- net_resnet_v1_beta.py:352-363
RESNET50_BLOCKS = [('block1', 64, 3, 2), ('block2', 128, 4, 2), ('block3', 256, 6, 2), ('block4', 512, 3, 1)]
BACKBONE_DEPTH = {'xception_65': 2048, 'resnet_v1_50_beta': 2048}
SKIP_DEPTH = {'xception_65': 256, 'resnet_v1_50_beta': 256}


def head_channels(num_objs, num_frags):
    """common.get_outputs_to_num_channels (/root/reference/epos_lib/common.py:189-203)."""
    return {
        'pred_obj_conf': num_objs + 1,
        'pred_frag_conf': num_objs * num_frags,
        'pred_frag_loc': num_objs * num_frags * 3,
    }


def variable_specs(num_objs, num_frags, model_variant='xception_65'):
    """List of (name, shape, init, stddev) for the backbone + ASPP + decoder + logits.

    `init` in {'tn' (truncated normal), 'vs' (slim.variance_scaling_initializer: truncated normal with
    stddev sqrt(1.3 * 2 / fan_in), external/slim/nets/resnet_utils.py:263), 'xavier', 'zeros', 'bn'}.
    For 'bn' the name is the BatchNorm scope and four variables are created under it.
    """
    specs = []

    def conv(scope, kh, cin, cout, init='tn', std=0.09, bn=True):
        specs.append((scope + '/weights', (kh, kh, cin, cout), init, std))
        if bn:
            specs.append((scope + '/BatchNorm', (cout,), 'bn', 0.0))

    def dw(scope, c, std=0.09):
        specs.append((scope + '/depthwise_weights', (3, 3, c, 1), 'tn', std))
        specs.append((scope + '/BatchNorm', (c,), 'bn', 0.0))

    if model_variant == 'xception_65':
        conv(XC + '/entry_flow/conv1_1', 3, 3, 32)
        conv(XC + '/entry_flow/conv1_2', 3, 32, 64)
        cin = 64
        for scope, depths, skip, units in XCEPTION65_BLOCKS:
            for u in range(1, units + 1):
                base = '%s/%s/unit_%d/xception_module' % (XC, scope, u)
                c = cin
                for i, d in enumerate(depths):
                    dw('%s/separable_conv%d_depthwise' % (base, i + 1), c)
                    conv('%s/separable_conv%d_pointwise' % (base, i + 1), 1, c, d)
                    c = d
                if skip == 'conv':
                    conv(base + '/shortcut', 1, cin, depths[-1])
                cin = depths[-1]
    elif model_variant == 'resnet_v1_50_beta':
        # root of three 3x3 convs (net_resnet_v1_beta.py:96-112), then bottleneck units (:38-93)
        conv(RN + '/conv1_1', 3, 3, 64, init='vs')
        conv(RN + '/conv1_2', 3, 64, 64, init='vs')
        conv(RN + '/conv1_3', 3, 64, 128, init='vs')
        cin = 128
        for scope, base_depth, units, _ in RESNET50_BLOCKS:
            for u in range(1, units + 1):
                base = '%s/%s/unit_%d/bottleneck_v1' % (RN, scope, u)
                if cin != base_depth * 4:
                    conv(base + '/shortcut', 1, cin, base_depth * 4, init='vs')
                conv(base + '/conv1', 1, cin, base_depth, init='vs')
                conv(base + '/conv2', 3, base_depth, base_depth, init='vs')
                conv(base + '/conv3', 1, base_depth, base_depth * 4, init='vs')
                cin = base_depth * 4
    else:
        raise ValueError('unsupported model_variant %r' % (model_variant,))
    # ASPP (model.py:217-258)
    conv('image_pooling', 1, 2048, 256, init='xavier')
    conv('aspp0', 1, 2048, 256, init='xavier')
    for i in (1, 2, 3):
        dw('aspp%d_depthwise' % i, 2048, std=0.33)
        conv('aspp%d_pointwise' % i, 1, 2048, 256, std=0.06)
    conv('concat_projection', 1, 1280, 256, init='xavier')
    # Decoder (model.py:325-380)
    conv('decoder/feature_projection0', 1, 256, 48, init='xavier')
    dw('decoder/decoder_conv0_depthwise', 304, std=0.33)
    conv('decoder/decoder_conv0_pointwise', 1, 304, 256, std=0.06)
    dw('decoder/decoder_conv1_depthwise', 256, std=0.33)
    conv('decoder/decoder_conv1_pointwise', 1, 256, 256, std=0.06)
    # Logit heads (model.py:432-456)
    for name, ch in sorted(head_channels(num_objs, num_frags).items()):
        conv('logits/' + name, 1, 256, ch, std=0.01, bn=False)
        specs.append(('logits/%s/biases' % name, (ch,), 'zeros', 0.0))
    return specs


def _truncated_normal(rng, shape, std):
    """tf.truncated_normal: values outside +-2 sigma are re-drawn."""
    x = rng.standard_normal(size=shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


def random_init(num_objs, num_frags, seed=0, bn='init', logits_std=None, model_variant='xception_65'):
    """Synthetic weights with the reference's initialisers.

    bn='init'   : gamma=1, beta=0, mean=0, var=1 (what a freshly initialised TF graph holds).
    bn='random' : perturbed statistics (gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1),
                  var~U(.5,1.5)); used by tests so that BN folding is actually exercised.
    logits_std  : override of the logit initialiser stddev (reference: 0.01).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}
    for name, shape, init, std in variable_specs(num_objs, num_frags, model_variant):
        if init == 'vs':
            kh, kw, cin, cout = shape
            w[name] = _truncated_normal(rng, shape, np.sqrt(1.3 * 2.0 / (kh * kw * cin)))
        elif init == 'tn':
            if logits_std is not None and name.startswith('logits/'):
                std = logits_std
            w[name] = _truncated_normal(rng, shape, std)
        elif init == 'xavier':
            kh, kw, cin, cout = shape
            lim = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
            w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif init == 'zeros':
            w[name] = np.zeros(shape, np.float32)
        elif init == 'bn':
            c = shape[0]
            if bn == 'init':
                vals = (np.ones(c), np.zeros(c), np.zeros(c), np.ones(c))
            else:
                vals = (rng.uniform(0.5, 1.5, c), rng.normal(0, 0.1, c),
                        rng.normal(0, 0.1, c), rng.uniform(0.5, 1.5, c))
            for k, v in zip(BN_KEYS, vals):
                w['%s/%s' % (name, k)] = v.astype(np.float32)
        else:
            raise ValueError(init)
    return w


def save_npz(path, weights):
    np.savez(path, **{k.replace('/', '|'): v for k, v in weights.items()})


def load_npz(path):
    with np.load(path) as z:
        return {k.replace('|', '/'): z[k] for k in z.files}


def synthetic_images(batch, seed=0, height=480, width=640):
    """uint8-valued f32 images in [0,255], NHWC."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, size=(batch, height, width, 3)).astype(np.float32)
