"""Input side (resize to max_height_before_crop + crop + intrinsics; datagen.py:424-476): oracle properties on the CPU,
device kernel against the oracle on the GPU."""
import numpy as np
import pytest

from oracle import preprocess as opre

K0 = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]])


def test_oracle_sizes_and_intrinsics():
    assert opre.new_size(480, 640, 480)[:2] == (480, 640)
    assert opre.new_size(1080, 1920, 480)[:2] == (480, 853)            # width truncated as tf.cast(int32) does
    assert opre.new_size(360, 480, 480)[:2] == (360, 480)              # never enlarged
    img = np.random.default_rng(0).integers(0, 256, (960, 1440, 3), dtype=np.uint8)
    out, K = opre.preprocess(img, K0, 480, (640, 480), offset=(0, 40))
    assert out.shape == (480, 640, 3) and out.dtype == np.float32
    assert np.isclose(K[0, 0], K0[0, 0] * 0.5) and np.isclose(K[0, 2], K0[0, 2] * 0.5 - 40) and np.isclose(K[1, 2], K0[1, 2] * 0.5)


def test_oracle_resize_area_properties():
    rng = np.random.default_rng(1)
    const = np.full((40, 60, 3), 77, np.uint8)
    assert np.allclose(opre.resize_area(const, 20, 30), 77.0, atol=1e-4)                 # weights sum to one
    img = rng.integers(0, 256, (24, 36, 3), dtype=np.uint8)
    assert np.allclose(opre.resize_area(img, 24, 36), img.astype(np.float32), atol=1e-4)  # identity at equal size
    # a smooth ramp stays a ramp (area averaging is linear): value at output y is the mean over its input span
    ramp = np.tile(np.arange(64, dtype=np.uint8)[:, None, None], (1, 8, 3))
    r = opre.resize_area(ramp, 16, 8)
    s = np.float32(63) / np.float32(15)
    assert np.allclose(r[5, 0, 0], (5 * s + 6 * s) / 2 - 0.5, atol=0.6) and np.all(np.diff(r[:, 0, 0]) > 0)


def test_oracle_bilinear_branch_matches_torch_align_corners():
    import torch
    import torch.nn.functional as F
    img = np.random.default_rng(2).integers(0, 256, (30, 40, 3), dtype=np.uint8)
    got = opre.resize_bilinear(img, 48, 64)
    ref = F.interpolate(torch.from_numpy(img.astype(np.float32)).permute(2, 0, 1)[None], size=(48, 64), mode='bilinear',
                        align_corners=True)[0].permute(1, 2, 0).numpy()
    assert np.abs(got - ref).max() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize('H,W,maxh,crop,off', [(480, 640, 480, (640, 480), (0, 0)), (960, 1440, 480, (640, 480), (0, 37)),
                                               (1080, 1920, 480, (640, 480), (0, 213)), (600, 800, 480, (640, 480), (0, 0)),
                                               (96, 128, 480, (100, 90), (3, 7)), (500, 700, 2000, (640, 480), (11, 29))])
def test_device_preprocess_matches_oracle(H, W, maxh, crop, off):
    import torch
    from epos_b200 import preprocess
    img = np.random.default_rng(H + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref, Kr = opre.preprocess(img, K0, maxh, crop, off)
    out, K = preprocess.prepare_image(torch.from_numpy(img).cuda(), K0, maxh, crop, off)
    torch.cuda.synchronize()
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-3                # f32 sums of up to ~9 taps of [0,255] values
    assert np.array_equal(K, Kr)
    with pytest.raises(RuntimeError):
        preprocess.prepare_image(torch.from_numpy(img).cuda(), K0, maxh, crop, (10 ** 6, 0))


@pytest.mark.gpu
def test_prepared_image_feeds_the_network():
    import torch
    from epos_b200 import model, preprocess, weights as Wt
    img = np.random.default_rng(5).integers(0, 256, (200, 300, 3), dtype=np.uint8)
    out, K = preprocess.prepare_image(torch.from_numpy(img).cuda(), K0, 96, (128, 96), None, np.random.default_rng(1))
    O, F = 2, 8
    w = Wt.random_init(O, F, seed=3)
    net = model.EposNet(w, O, F, 'cuda:0', model_options=model.ModelOptions(Wt.head_channels(O, F), crop_size=(128, 96)))
    pred = net.predict(out[None])
    torch.cuda.synchronize()
    assert pred[model.PRED_OBJ_CONF].shape == (1, 24, 32, O + 1) and bool(torch.isfinite(pred[model.PRED_FRAG_LOC]).all())
