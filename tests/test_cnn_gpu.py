"""GPU parity tests of the CNN kernels against the oracle (oracle/cnn.py), through the C ABI.

Tolerance (north_star): output maps within 1e-3 relative (to the per-tensor max-abs) of the f32 oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='module')
def env():
    from epos_b200 import _lib
    return _lib.lib(), torch.device('cuda:0')


def split(t):
    hi = t.to(torch.bfloat16)
    return torch.stack([hi, (t - hi.float()).to(torch.bfloat16)]).contiguous()


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (300, 256, 128), (4800, 728, 728), (1000, 22, 256),
                                   (777, 48, 256), (2500, 1344, 256), (600, 256, 1280), (384, 1024, 2048),
                                   (129, 304, 304)])
@pytest.mark.parametrize('mode', ['plain', 'relu_res', 'split_out'])
def test_pw_gemm(env, M, N, K, mode):
    from epos_b200 import _lib
    lib, dev = env
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.1).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev) if mode == 'relu_res' else None
    relu = mode != 'plain'
    ref = a.double() @ w.double().T + bias.double()
    if res is not None:
        ref = ref + res.double()
    if relu:
        ref = ref.clamp_min(0)                       # ReLU follows the residual add (net_resnet_v1_beta.py:88)
    a_s, w_s = split(a), split(w)
    d = torch.full((M, N), float('nan'), device=dev) if mode != 'split_out' else None
    ds = torch.zeros((2, M, N), dtype=torch.bfloat16, device=dev) if mode == 'split_out' else None
    rc = lib.epos_pwconv_gemm(a_s.data_ptr(), K, a_s.stride(0), w_s.data_ptr(), K, bias.data_ptr(), 0,
                              _lib.ptr(res), N, _lib.ptr(d), N, _lib.ptr(ds), N, 0 if ds is None else ds.stride(0),
                              M, N, K, int(relu), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, 'gemm')
    torch.cuda.synchronize()
    got = d if d is not None else ds[0].float() + ds[1].float()
    tol = 2e-5 if d is not None else 5e-5
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < tol


@pytest.mark.parametrize('M,N,K', [(300, 128, 64), (4000, 1344, 256), (129, 64, 40)])
def test_pw_gemm_fused_softmax64(env, M, N, K):
    """relu = 2: softmax over aligned groups of 64 output columns in the GEMM epilogue (model.py:676-678)."""
    from epos_b200 import _lib
    lib, dev = env
    g = torch.Generator(device='cpu').manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ref = torch.softmax((a.double() @ w.double().T + bias.double()).view(M, N // 64, 64), dim=-1).view(M, N)
    a_s, w_s = split(a), split(w)
    d = torch.full((M, N), float('nan'), device=dev)
    rc = lib.epos_pwconv_gemm(a_s.data_ptr(), K, a_s.stride(0), w_s.data_ptr(), K, bias.data_ptr(), 0, None, 0,
                              d.data_ptr(), N, None, 0, 0, M, N, K, 2, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, 'gemm')
    torch.cuda.synchronize()
    assert rel_err(d.cpu().numpy(), ref.cpu().numpy()) < 1e-4
    assert abs(float(d.sum()) - M * N // 64) < 1e-2 * M


def test_pw_gemm_balanced_pieces_cover_everything(env):
    """Shapes whose (m-tile, 8-column) units do not divide by the SM count: every output must be written exactly once
    (NaN-prefilled destination) and rows/columns outside [M, N) must stay untouched."""
    from epos_b200 import _lib
    lib, dev = env
    for M, N, K in [(38400, 728, 72), (4800, 728, 40), (1000, 1000, 24), (129, 264, 16), (5000, 48, 32)]:
        a = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) * 0.1
        a_s, w_s = split(a), split(w)
        buf = torch.full((M + 3, N + 8), float('nan'), device=dev)
        rc = lib.epos_pwconv_gemm(a_s.data_ptr(), K, a_s.stride(0), w_s.data_ptr(), K, None, 0, None, 0,
                                  buf.data_ptr(), N + 8, None, 0, 0, M, N, K, 0, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'gemm')
        torch.cuda.synchronize()
        ref = a.double() @ w.double().T
        assert rel_err(buf[:M, :N].cpu().numpy(), ref.cpu().numpy()) < 2e-5, (M, N, K)
        assert bool(torch.isnan(buf[M:]).all()) and bool(torch.isnan(buf[:, N:]).all()), (M, N, K)


@pytest.mark.parametrize('M,N,K,reps', [(153600, 256, 64, 12), (38400, 728, 128, 4), (9000, 512, 200, 4)])
def test_pw_gemm_pair_three_outputs_repeated(env, M, N, K, reps):
    """CTA-pair GEMM with the slowest epilogue (residual + f32 + split-bf16 outputs) on short-K shapes, launched
    repeatedly: the case in which a consumer's "slot consumed" signal once overtook its read of the piece queue
    (the kernel then hung on 100 % of the launches of 153600 x 256 x 64).  Every launch must finish and be exact."""
    from epos_b200 import _lib
    lib, dev = env
    g = torch.Generator(device='cpu').manual_seed(N + K)
    a = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.1).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev)
    ref = (a[:4096].double() @ w.double().T + bias.double() + res[:4096].double()).clamp_min(0)
    a_s, w_s = split(a), split(w)
    s = torch.cuda.current_stream().cuda_stream
    for it in range(reps):
        d = torch.full((M, N), float('nan'), device=dev)
        ds = torch.zeros((2, M, N), dtype=torch.bfloat16, device=dev)
        _lib.check(lib.epos_pwconv_gemm(a_s.data_ptr(), K, a_s.stride(0), w_s.data_ptr(), K, bias.data_ptr(), 0,
                                        res.data_ptr(), N, d.data_ptr(), N, ds.data_ptr(), N, ds.stride(0),
                                        M, N, K, 1, s), 'gemm')
        torch.cuda.synchronize()
        assert not bool(torch.isnan(d).any()), it
        assert rel_err(d[:4096].cpu().numpy(), ref.cpu().numpy()) < 2e-5, it
        assert torch.equal(ds[0], d.to(torch.bfloat16)), it
        assert float((ds[0][-4096:].float() + ds[1][-4096:].float() - d[-4096:]).abs().max()) <= 2e-5 * float(d.abs().max())


def test_pw_gemm_grouped_bias_and_slices(env):
    """per-image bias rows (image-pooling fold) and strided output slices (concat buffers)."""
    from epos_b200 import _lib
    lib, dev = env
    M, N, K, G = 960, 48, 256, 4
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    bias = torch.randn(G, N, device=dev)
    buf = torch.zeros(M, 304, device=dev)
    a_s, w_s = split(a), split(w)
    rc = lib.epos_pwconv_gemm(a_s.data_ptr(), K, a_s.stride(0), w_s.data_ptr(), K, bias.data_ptr(), M // G, None, 0,
                              buf[:, 256:].data_ptr(), 304, None, 0, 0, M, N, K, 1,
                              torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, 'gemm')
    ref = (a.double() @ w.double().T + bias.double().repeat_interleave(M // G, 0)).clamp_min(0)
    assert rel_err(buf[:, 256:].cpu().numpy(), ref.cpu().numpy()) < 2e-5
    assert float(buf[:, :256].abs().max()) == 0.0


@pytest.mark.parametrize('C,H,W,stride,rate,relu_in,relu_out', [
    (64, 24, 32, 1, 1, True, False), (128, 24, 32, 2, 1, True, False), (728, 15, 20, 1, 2, True, False),
    (1024, 15, 20, 1, 4, False, True), (2048, 15, 20, 1, 12, False, True), (304, 30, 40, 1, 1, False, True),
    (256, 9, 11, 2, 1, True, False), (2048, 15, 20, 1, 36, False, True),
    (728, 60, 80, 1, 2, True, False), (132, 50, 33, 1, 4, True, True), (20, 17, 19, 1, 1, False, False),
    (2048, 60, 80, 1, 24, True, True), (100, 60, 80, 1, 12, True, False), (64, 31, 45, 1, 6, False, False)])
def test_dwconv(env, C, H, W, stride, rate, relu_in, relu_out):
    from epos_b200 import _lib
    from oracle import cnn
    lib, dev = env
    B = 2
    rng = np.random.default_rng(C + rate)
    x = rng.standard_normal((B, H, W, C)).astype(np.float32)
    k = rng.standard_normal((3, 3, C, 1)).astype(np.float32)
    b = rng.standard_normal(C).astype(np.float32)
    o = cnn.Oracle({'s/depthwise_weights': k})
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    if relu_in:
        xt = xt.clamp_min(0)
    if stride == 1:
        y = o.depthwise(xt, 's', 1, rate, 'SAME')
    else:
        y = o.depthwise(o.fixed_padding(xt, 3, rate), 's', stride, rate, 'VALID')
    y = y + torch.from_numpy(b).view(1, -1, 1, 1)
    if relu_out:
        y = y.clamp_min(0)
    ref = y.permute(0, 2, 3, 1).numpy()
    Ho, Wo = ref.shape[1:3]
    xd = torch.from_numpy(x).to(dev)
    wd = torch.from_numpy(k[:, :, :, 0].reshape(9, C).copy()).to(dev)
    bd = torch.from_numpy(b).to(dev)
    y32 = torch.empty((B * Ho * Wo, C), device=dev)
    ldy = (C + 15) // 16 * 16                      # padded bf16 row pitch, as model.py allocates it
    ys = torch.full((2, B * Ho * Wo, ldy), 7.0, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.epos_dwconv3x3(xd.data_ptr(), C, wd.data_ptr(), bd.data_ptr(), y32.data_ptr(), ys.data_ptr(), ldy, B, H, W,
                                  C, stride, rate, int(relu_in), int(relu_out),
                                  torch.cuda.current_stream().cuda_stream), 'dw')
    torch.cuda.synchronize()
    assert rel_err(y32.cpu().numpy().reshape(ref.shape), ref) < 1e-5
    assert rel_err((ys[0, :, :C].float() + ys[1, :, :C].float()).cpu().numpy().reshape(ref.shape), ref) < 3e-5
    assert bool((ys[:, :, C:] == 7.0).all())       # the pad columns are never written


@pytest.mark.parametrize('B,H,W,C,N,rate,mode', [
    (2, 16, 32, 64, 64, 1, 'relu'), (1, 15, 20, 128, 128, 2, 'relu'), (2, 17, 19, 64, 40, 1, 'plain'),
    (1, 30, 40, 256, 256, 4, 'res'), (1, 12, 16, 512, 512, 8, 'relu'), (3, 8, 16, 64, 128, 1, 'split')])
def test_conv3x3_gemm(env, B, H, W, C, N, rate, mode):
    """Implicit-GEMM 3x3 atrous conv (TMA-shifted taps) against F.conv2d with TF SAME padding."""
    from epos_b200 import _lib
    from oracle import cnn
    lib, dev = env
    rng = np.random.default_rng(H * W + C + N + rate)
    x = rng.standard_normal((B, H, W, C)).astype(np.float32)
    k = (rng.standard_normal((3, 3, C, N)) * 0.05).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    res = rng.standard_normal((B, H, W, N)).astype(np.float32) if mode == 'res' else None
    o = cnn.Oracle({'s/weights': k}, dtype=torch.float64)
    y = o.conv(torch.from_numpy(x).double().permute(0, 3, 1, 2), 's', 1, rate, 'SAME').permute(0, 2, 3, 1)
    y = y + torch.from_numpy(bias).double()
    if res is not None:
        y = y + torch.from_numpy(res).double()
    if mode != 'plain':
        y = y.clamp_min(0)
    ref = y.numpy().reshape(B * H * W, N)
    xs = split(torch.from_numpy(x).to(dev).view(B * H * W, C))
    ws = split(torch.from_numpy(k.reshape(9 * C, N).T.copy()).to(dev))          # [N][9C], k = (ky*3+kx)*C + c
    bd = torch.from_numpy(bias).to(dev)
    rd = torch.from_numpy(res).to(dev).view(B * H * W, N) if res is not None else None
    d = torch.full((B * H * W, N), float('nan'), device=dev) if mode != 'split' else None
    ds = torch.zeros((2, B * H * W, N), dtype=torch.bfloat16, device=dev) if mode == 'split' else None
    rc = lib.epos_conv3x3_gemm(xs.data_ptr(), C, xs.stride(0), ws.data_ptr(), 9 * C, bd.data_ptr(), _lib.ptr(rd), N,
                               _lib.ptr(d), N, _lib.ptr(ds), N, 0 if ds is None else ds.stride(0), B, H, W, C, N, rate,
                               int(mode != 'plain'), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, 'conv3x3_gemm')
    torch.cuda.synchronize()
    got = d if d is not None else ds[0].float() + ds[1].float()
    assert rel_err(got.cpu().numpy(), ref) < (2e-5 if d is not None else 5e-5)


@pytest.mark.parametrize('H,W', [(16, 24), (15, 21), (9, 8)])
def test_maxpool_and_subsample(env, H, W):
    from epos_b200 import _lib
    lib, dev = env
    B, C = 2, 72
    x = torch.randn(B, H, W, C, device=dev)
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    ph, pw = max((Ho - 1) * 2 + 3 - H, 0), max((Wo - 1) * 2 + 3 - W, 0)
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float('-inf'))
    ref = torch.nn.functional.max_pool2d(xp, 3, 2).permute(0, 2, 3, 1)
    y = torch.empty(B, Ho, Wo, C, device=dev)
    ys = torch.empty(2, B * Ho * Wo, C, dtype=torch.bfloat16, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.epos_maxpool3x3_s2(x.data_ptr(), y.data_ptr(), ys.data_ptr(), B, H, W, C, s), 'maxpool')
    z = torch.empty(B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C, device=dev)
    _lib.check(lib.epos_subsample_f32(x.data_ptr(), C, z.data_ptr(), B, H, W, C, 2, s), 'subsample')
    torch.cuda.synchronize()
    assert torch.equal(y, ref)
    assert rel_err((ys[0].float() + ys[1].float()).view(B, Ho, Wo, C).cpu().numpy(), ref.cpu().numpy()) < 3e-5
    assert torch.equal(z, x[:, ::2, ::2])


def test_entry_convs(env):
    from epos_b200 import _lib, weights as W
    from oracle import cnn
    lib, dev = env
    w = {k: v for k, v in W.random_init(1, 1, seed=5, bn='random').items() if 'entry_flow/conv1_' in k}
    img = W.synthetic_images(2, seed=5, height=64, width=96)
    o = cnn.Oracle(w)
    x = torch.from_numpy(img).permute(0, 3, 1, 2)
    x = (2.0 / 255.0) * x - 1.0
    r1 = o.conv2d_same(x, 'xception_65/entry_flow/conv1_1', 2)
    r2 = o.conv2d_same(r1, 'xception_65/entry_flow/conv1_2', 1)
    from epos_b200.model import EposNet, _bn_fold
    def fold(scope):
        k = np.asarray(w[scope + '/weights'], np.float64)
        s, sh = _bn_fold(w, scope, 1e-3)
        return (torch.from_numpy((k * s).astype(np.float32)).to(dev), torch.from_numpy(sh.astype(np.float32)).to(dev))
    k1, b1 = fold('xception_65/entry_flow/conv1_1')
    k2, b2 = fold('xception_65/entry_flow/conv1_2')
    xd = torch.from_numpy(img).to(dev)
    c1 = torch.empty((2, 32, 48, 32), device=dev)
    c2 = torch.empty((2, 32, 48, 64), device=dev)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.epos_conv3x3_rgb_s2(xd.data_ptr(), k1.data_ptr(), b1.data_ptr(), c1.data_ptr(), None, 2, 64, 96, 32, s), 'c1')
    _lib.check(lib.epos_conv3x3_dense(c1.data_ptr(), k2.data_ptr(), b2.data_ptr(), c2.data_ptr(), 2, 32, 48, 32, 64, s), 'c2')
    torch.cuda.synchronize()
    assert rel_err(c1.cpu().numpy(), r1.permute(0, 2, 3, 1).numpy()) < 1e-5
    assert rel_err(c2.cpu().numpy(), r2.permute(0, 2, 3, 1).numpy()) < 1e-5


@pytest.mark.parametrize('cout,H,W', [(32, 64, 96), (64, 37, 51), (32, 480, 640)])
def test_stem_conv_split_output(env, cout, H, W):
    """conv1_1 (3x3 stride 2 on the raw image, preprocessing fused): f32 and split-bf16 outputs, both widths, odd sizes."""
    from epos_b200 import _lib, weights as Wt
    import torch.nn.functional as F
    lib, dev = env
    B = 2
    img = torch.from_numpy(Wt.synthetic_images(B, seed=H + cout, height=H, width=W)).to(dev)
    g = torch.Generator(device='cpu').manual_seed(cout)
    k = (torch.randn(3, 3, 3, cout, generator=g) * 0.3).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(B, Ho, Wo, cout, device=dev)
    ys = torch.empty(2, B, Ho, Wo, cout, dtype=torch.bfloat16, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.epos_conv3x3_rgb_s2(img.data_ptr(), k.data_ptr(), b.data_ptr(), y.data_ptr(), ys.data_ptr(), B, H, W, cout, s), 'c1')
    torch.cuda.synchronize()
    x = ((2.0 / 255.0) * img.double() - 1.0).permute(0, 3, 1, 2)
    ref = F.relu(F.conv2d(F.pad(x, (1, 1, 1, 1)), k.double().permute(3, 2, 0, 1), b.double(), stride=2))
    ref = ref.permute(0, 2, 3, 1)[:, :Ho, :Wo]
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    rec = ys[0].float() + ys[1].float()
    assert float((rec - y).abs().max()) <= 2e-5 * float(y.abs().max())
    assert torch.equal(ys[0], y.to(torch.bfloat16))


def test_resize_mean_softmax(env):
    from epos_b200 import _lib
    import torch.nn.functional as F
    lib, dev = env
    s = torch.cuda.current_stream().cuda_stream
    x = torch.randn(2, 15, 20, 256, device=dev)
    y = torch.zeros(2, 30, 40, 304, device=dev)
    _lib.check(lib.epos_resize_bilinear(x.data_ptr(), y.data_ptr(), 304, 2, 15, 20, 30, 40, 256, s), 'rs')
    ref = F.interpolate(x.permute(0, 3, 1, 2), size=(30, 40), mode='bilinear', align_corners=True).permute(0, 2, 3, 1)
    assert rel_err(y[..., :256].cpu().numpy(), ref.cpu().numpy()) < 1e-5
    assert float(y[..., 256:].abs().max()) == 0
    m = torch.empty(2, 256, device=dev)
    _lib.check(lib.epos_global_mean(x.data_ptr(), m.data_ptr(), 2, 300, 256, s), 'mean')
    assert rel_err(m.cpu().numpy(), x.view(2, 300, 256).mean(1).cpu().numpy()) < 1e-5
    for n in (2, 22, 64, 256):
        z = torch.randn(1000, n, device=dev) * 3
        ref = torch.softmax(z, -1)
        lab = torch.empty(1000, dtype=torch.int64, device=dev)
        _lib.check(lib.epos_softmax_rows(z.data_ptr(), lab.data_ptr(), 1000, n, s), 'sm')
        assert rel_err(z.cpu().numpy(), ref.cpu().numpy()) < 1e-5
        assert (lab == ref.argmax(-1)).float().mean() > 0.999


def test_softmax_rows_masked(env):
    """Engine path: the fragment softmax only where the object confidence passes -- bit-identical to the full row softmax
    on those rows, logits untouched elsewhere."""
    from epos_b200 import _lib
    lib, dev = env
    s = torch.cuda.current_stream().cuda_stream
    for P, O, F, thr in ((5000, 30, 256, 0.1), (777, 3, 128, 0.3), (64, 21, 20, 0.0)):
        g = torch.Generator(device='cpu').manual_seed(P + O)
        x = (torch.randn(P, O, F, generator=g) * 4).to(dev)
        oc = torch.softmax(torch.randn(P, O + 1, generator=g) * 3, -1).to(dev).contiguous()
        full = x.clone()
        _lib.check(lib.epos_softmax_rows(full.data_ptr(), None, P * O, F, s), 'sm')
        y = x.clone()
        _lib.check(lib.epos_softmax_rows_masked(y.data_ptr(), oc.data_ptr(), P, O, F, float(thr), s), 'smm')
        torch.cuda.synchronize()
        mask = oc[:, 1:] > np.float32(thr)
        assert 0 < int(mask.sum()) < P * O or thr == 0.0
        assert torch.equal(y[mask], full[mask])
        assert torch.equal(y[~mask], x[~mask])


def _net_parity(B, H, W, O, F, seed, check_simt=False, variant='xception_65', multi_grid=None, logits_std=0.5):
    from epos_b200 import model, weights as Wt
    from oracle import cnn
    dev = torch.device('cuda:0')
    w = Wt.random_init(O, F, seed=seed, bn='random', logits_std=logits_std, model_variant=variant)
    img = Wt.synthetic_images(B, seed=seed, height=H, width=W)
    opts = model.ModelOptions(Wt.head_channels(O, F), crop_size=(W, H), model_variant=variant, multi_grid=multi_grid)
    net = model.EposNet(w, O, F, dev, model_options=opts, keep_f32=True)
    out = net.predict(torch.from_numpy(img).to(dev))
    torch.cuda.synchronize()
    ref = cnn.predict(w, img, O, F, return_features=True, model_variant=variant, multi_grid=multi_grid)
    errs = {}
    ep = model.RESNET_END_POINT if variant == 'resnet_v1_50_beta' else model.DECODER_END_POINT
    net.end_points['skip'] = net.end_points[ep]
    for name, key in (('skip', '_skip'), ('backbone', '_backbone'), ('aspp', '_aspp'), ('decoder', '_decoder')):
        t, h, w_, c = net.end_points[name]
        errs[name] = rel_err(t.cpu().numpy().reshape(ref[key].shape), ref[key])
    for k in (model.PRED_OBJ_CONF, model.PRED_FRAG_CONF, model.PRED_FRAG_LOC):
        assert out[k].shape == ref[k].shape and out[k].dtype == torch.float32
        errs[k] = rel_err(out[k].cpu().numpy(), ref[k])
    lab = out[model.PRED_OBJ_LABEL]
    assert lab.dtype == torch.int64 and lab.shape == ref['pred_obj_label'].shape
    errs['label_agree'] = float((lab.cpu().numpy() == ref['pred_obj_label']).mean())
    print(errs)
    for k, v in errs.items():
        if k == 'label_agree':
            assert v > 0.999
        else:
            assert v < 1e-3, (k, v)
    if check_simt:
        net.impl = 'simt'
        out2 = net.predict(torch.from_numpy(img).to(dev))
        for k in (model.PRED_FRAG_LOC, model.PRED_OBJ_CONF):
            assert rel_err(out2[k].cpu().numpy(), ref[k]) < 1e-3


def test_network_small():
    _net_parity(2, 96, 128, 3, 8, seed=11, check_simt=True)


def test_network_fragment_counts_around_the_fused_softmax_width():
    """F = 64 takes the softmax fused into the GEMM epilogue; F = 128 (a multiple of 64, like config 5's 256) and F = 32 must
    take the row-softmax kernel (softmax over ALL fragments of an object)."""
    _net_parity(1, 64, 96, 2, 128, seed=31)
    _net_parity(1, 64, 96, 2, 64, seed=32)
    _net_parity(1, 64, 96, 3, 32, seed=33)


def test_network_odd_size():
    _net_parity(1, 81, 113, 1, 4, seed=12)


def test_network_full_size_c1():
    """BASELINE config 1: single 640x480 image, 1 object / 64 fragments."""
    _net_parity(1, 480, 640, 1, 64, seed=13)


def test_network_full_size_ycbv_heads():
    """21 objects x 64 fragments (YCB-V-shaped heads), batch 2."""
    _net_parity(2, 480, 640, 21, 64, seed=14)


def test_network_full_size_config5_heads():
    """BASELINE configs[4] head shape: 30 objects x 256 fragments at 640x480 (N = 7 680 conf + 23 040 loc columns,
    2.36 GB of head maps per image, row-softmax path for F = 256)."""
    _net_parity(1, 480, 640, 30, 256, seed=15)


# The variance-scaling initialiser with perturbed BN statistics leaves decoder features of magnitude ~1e3-1e4; the logit
# stddev is scaled down so that the logits stay O(1-10) and the softmax outputs are a meaningful comparison.
def test_resnet50_beta_small():
    """BASELINE config 4 backbone (resnet_v1_50_beta, net_resnet_v1_beta.py:302-373) at a small size."""
    _net_parity(2, 96, 128, 3, 8, seed=21, variant='resnet_v1_50_beta', logits_std=0.002)


def test_resnet50_beta_odd_size_multigrid():
    _net_parity(1, 81, 113, 2, 4, seed=22, variant='resnet_v1_50_beta', multi_grid=(1, 2, 4), logits_std=0.002)


def test_resnet50_beta_full_size():
    _net_parity(1, 480, 640, 21, 64, seed=23, variant='resnet_v1_50_beta', logits_std=0.0005)
