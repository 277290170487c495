"""Pins the oracle's conv/padding restatement on the reference's only CNN golden vectors:
TF-Slim ResnetUtilsTest (/root/reference/external/slim/nets/resnet_v1_test.py:58-149)."""
import numpy as np
import torch

from oracle import cnn


def mesh(n, m):
    return (np.arange(n).reshape(n, 1) + np.arange(m).reshape(1, m)).astype(np.float32)


def _oracle_with_kernel():
    w = {'Conv/weights': mesh(3, 3).reshape(3, 3, 1, 1)}
    return cnn.Oracle(w)


def _run(fn, x, **kw):
    xt = torch.from_numpy(x).reshape(1, 1, *x.shape)
    return fn(xt, 'Conv', **kw)[0, 0].numpy()


def test_conv2d_same_even_and_odd(golden):
    g = golden('slim_conv2d_same.json')
    o = _oracle_with_kernel()
    for key in ('testConv2DSameEven', 'testConv2DSameOdd'):
        n = g[key]['n']
        x = mesh(n, n)
        y1 = _run(o.conv, x, stride=1, padding='SAME')
        np.testing.assert_allclose(y1, np.array(g[key]['y1'], np.float32))
        # subsample(y1, 2)
        np.testing.assert_allclose(y1[::2, ::2], np.array(g[key]['y2'], np.float32))
        # conv2d_same(stride 2) == explicit padding then VALID
        y3 = _run(lambda t, s, **k: o.conv(o.fixed_padding(t, 3, 1), s, stride=2, padding='VALID'), x)
        np.testing.assert_allclose(y3, np.array(g[key]['y2'], np.float32))
        # plain SAME stride 2 differs on even sizes
        y4 = _run(o.conv, x, stride=2, padding='SAME')
        exp4 = g[key]['y4'] if g[key]['y4'] is not None else g[key]['y2']
        np.testing.assert_allclose(y4, np.array(exp4, np.float32))


def test_subsample_vectors(golden):
    g = golden('slim_conv2d_same.json')
    assert np.arange(9).reshape(3, 3)[::2, ::2].ravel().tolist() == g['subsample3']
    assert np.arange(16).reshape(4, 4)[::2, ::2].ravel().tolist() == g['subsample4']


def test_depthwise_matches_dense_on_single_channel():
    o = cnn.Oracle({'Conv/weights': mesh(3, 3).reshape(3, 3, 1, 1),
                    'Conv/depthwise_weights': mesh(3, 3).reshape(3, 3, 1, 1)})
    x = torch.from_numpy(mesh(6, 8)).reshape(1, 1, 6, 8)
    for rate in (1, 2):
        a = o.conv(x, 'Conv', 1, rate, 'SAME')
        b = o.depthwise(x, 'Conv', 1, rate, 'SAME')
        np.testing.assert_allclose(a.numpy(), b.numpy())


def test_atrous_equals_strided_subsample():
    """resnet_v1_test.py:197-240 idea: atrous conv at stride 1 sampled every 2nd pixel equals
    fixed_padding + stride-2 VALID conv."""
    o = _oracle_with_kernel()
    x = torch.randn(1, 1, 9, 9)
    a = o.conv(x, 'Conv', 1, 1, 'SAME')[:, :, ::2, ::2]
    b = o.conv(o.fixed_padding(x, 3, 1), 'Conv', 2, 1, 'VALID')
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-5)


def test_shapes_and_layout_small():
    """Channel order of the heads: c = o*F + f, loc c = (o*F+f)*3+k (model.py:133-147)."""
    from epos_b200 import weights as W
    w = W.random_init(2, 4, seed=1)
    img = W.synthetic_images(1, seed=1, height=64, width=96)
    out = cnn.predict(w, img, 2, 4)
    assert out['pred_obj_conf'].shape == (1, 16, 24, 3)
    assert out['pred_obj_label'].shape == (1, 16, 24) and out['pred_obj_label'].dtype == np.int64
    assert out['pred_frag_conf'].shape == (1, 16, 24, 2, 4)
    assert out['pred_frag_loc'].shape == (1, 16, 24, 2, 4, 3)
    np.testing.assert_allclose(out['pred_obj_conf'].sum(-1), 1.0, rtol=1e-5)
    np.testing.assert_allclose(out['pred_frag_conf'].sum(-1), 1.0, rtol=1e-5)


def test_resnet_beta_atrous_equals_nominal_stride_subsampled():
    """resnet_v1_test.py:470-495 (testAtrousFullyConvolutionalValues): dense features at output stride 8 sampled every
    4th pixel equal the nominal-stride-32 features (same weights); pins the stride -> atrous-rate conversion of the
    ResNet restatement (resnet_utils.py:125-217) and conv2d_same / subsample on an odd size."""
    from epos_b200 import weights as W
    w = W.random_init(1, 2, seed=3, bn='random', model_variant='resnet_v1_50_beta')
    img = W.synthetic_images(1, seed=3, height=65, width=97)
    o = cnn.Oracle(w, dtype=torch.float64, model_variant='resnet_v1_50_beta')
    with torch.no_grad():
        dense = o.resnet_backbone(img, output_stride=8)
        nominal = o.resnet_backbone(img, output_stride=32)
    assert dense.shape[2:] == (9, 13) and nominal.shape[2:] == (3, 4)
    np.testing.assert_allclose(dense[:, :, ::4, ::4].numpy(), nominal.numpy(), rtol=1e-8, atol=1e-8 * float(nominal.abs().max()))


def test_resnet_beta_shapes_and_end_point():
    from epos_b200 import weights as W
    w = W.random_init(2, 4, seed=1, model_variant='resnet_v1_50_beta')
    img = W.synthetic_images(1, seed=1, height=64, width=96)
    out = cnn.predict(w, img, 2, 4, model_variant='resnet_v1_50_beta', return_features=True)
    assert out['_backbone'].shape == (1, 8, 12, 2048) and out['_skip'].shape == (1, 16, 24, 256)
    assert out['pred_frag_loc'].shape == (1, 16, 24, 2, 4, 3)
