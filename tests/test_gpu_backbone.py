"""GPU suite for the conv5_3 backbone: image blob vs the reference's own _get_image_blob (golden)
and the oracle restatement; 2x2 ceil-mode max pooling bit-exact; the tcgen05 implicit-GEMM 3x3 convolution vs a
plain fp32 convolution of the same bf16-rounded operands; the whole VGG16 stack vs the fp32 oracle; and the
from-image proposal route."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from aznet_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu

BLOB_TOL = 6.2e-5       # 4 float32 ulps at |v| < 256 (cv2's SIMD path rounds its two multiply-adds differently)


@pytest.fixture(scope="module")
def dev():
    from aznet_b200 import _lib
    _lib.build()
    _lib.require_device()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def O():
    from oracle import az_oracle
    az_oracle.build()
    return az_oracle


def _bf(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


def test_image_blob_matches_reference_golden(dev, golden_blob):
    from aznet_b200 import ops
    from aznet_b200.backbone import PIXEL_MEANS
    names = sorted(k[:-3] for k in golden_blob.files if k.endswith("_im"))
    for name in names:
        im, ref, c = golden_blob[name + "_im"], golden_blob[name + "_blob"], golden_blob[name + "_cfg"]
        x = torch.from_numpy(im[None]).to(dev)
        padded, blob = ops.image_blob(x, float(c[2]), PIXEL_MEANS, 64, want_f32=True)
        blob = blob.cpu().numpy()
        assert blob.shape == ref.shape, (name, blob.shape, ref.shape)
        assert np.abs(blob - ref).max() <= BLOB_TOL, (name, np.abs(blob - ref).max())
        p = padded.float().cpu().numpy()
        hs, ws = ref.shape[2:]
        assert p.shape == (1, hs + 2, ws + 2, 64)
        assert np.array_equal(p[0, 1:-1, 1:-1, :3], _bf(blob[0].transpose(1, 2, 0)))     # interior = bf16 of the f32 blob
        assert not p[0, 1:-1, 1:-1, 3:].any()                                            # padded channels
        assert not p[0, 0].any() and not p[0, -1].any() and not p[0, :, 0].any() and not p[0, :, -1].any()   # zero border


def test_image_blob_batch_vs_oracle(dev, O):
    from aznet_b200 import ops
    from aznet_b200.backbone import PIXEL_MEANS
    rng = np.random.RandomState(3)
    ims = rng.randint(0, 256, (3, 120, 200, 3)).astype(np.uint8)
    cfg = O.OracleCfg(TEST_SCALES=(96,), TEST_MAX_SIZE=1000)
    s = O.im_scale_for(ims.shape[1:], cfg)[0]
    _, blob = ops.image_blob(torch.from_numpy(ims).to(dev), float(s), PIXEL_MEANS, 64, want_f32=True)
    for i in range(3):
        ref, s2 = O.get_image_blob(ims[i], cfg)
        assert s2 == s and np.abs(blob[i:i + 1].cpu().numpy() - ref).max() <= BLOB_TOL


@pytest.mark.parametrize("n,H,W,C", [(2, 13, 17, 64), (1, 30, 50, 128), (3, 8, 8, 8), (1, 1, 5, 64)])
def test_maxpool2x2_bit_exact(dev, n, H, W, C):
    from aznet_b200 import ops
    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn((n, H, W, C), generator=g).to(torch.bfloat16)
    ref = torch.nn.functional.max_pool2d(x.float().permute(0, 3, 1, 2), 2, 2, ceil_mode=True).permute(0, 2, 3, 1)
    xp = ops.nhwc_border(x.to(dev).contiguous(), True)
    # dirty output buffer: the kernel must write its own border
    out = torch.full((n, (H + 1) // 2 + 2, (W + 1) // 2 + 2, C), 7.0, dtype=torch.bfloat16, device=dev)
    y = ops.maxpool2x2(xp, out=out)
    got = ops.nhwc_border(y, False).float().cpu()
    assert torch.equal(got, ref)
    yb = y.float().cpu()
    assert not yb[:, 0].any() and not yb[:, -1].any() and not yb[:, :, 0].any() and not yb[:, :, -1].any()


@pytest.fixture(params=[(1, -1), (1, 128), (0, 0)], ids=["row_reuse", "row_reuse_wide256", "box_per_tap"])
def conv_variant(request):
    """azn_conv_tune: the RU kernels (one activation box per filter row, its three taps through row-shifted descriptors;
    wide layers on 256 x 128 tiles), the same with 256 x 256 tiles for wide layers of more than 128 input channels, and
    round 1's kernel (one box per tap)."""
    from aznet_b200 import _lib
    _lib.lib().azn_conv_tune(*request.param)
    yield request.param
    _lib.lib().azn_conv_tune(1, -1)


def test_conv3x3_row_reuse_same_bits_at_cin64(dev):
    """With one channel block per tap the RU kernel accumulates in the same order as the kernel with one box per tap
    (filter row, tap, 16-wide k step): identical output bits -- the row-shifted descriptors read exactly the rows a
    shifted box would hold (and a map wider than a tile, so that tiles start in the middle of image rows)."""
    from aznet_b200 import _lib, ops
    g = torch.Generator().manual_seed(77)
    outs = []
    for n, H, W, Cout in ((2, 37, 301, 64), (1, 50, 90, 128)):
        x = torch.randn((n, H, W, 64), generator=g).to(torch.bfloat16)
        w = (torch.randn((Cout, 64, 3, 3), generator=g) * (2.0 / (9 * 64)) ** 0.5).to(torch.bfloat16)
        b = torch.randn((Cout,), generator=g) * 0.1
        xp = ops.nhwc_border(x.to(dev).contiguous(), True)
        wt = ops.pack_conv_weight(w.float().to(dev))
        got = []
        try:
            for variant in ((1, -1), (0, 0)):
                _lib.lib().azn_conv_tune(*variant)
                got.append(ops.conv3x3(xp, wt, b.to(dev), relu=True).clone())
        finally:
            _lib.lib().azn_conv_tune(1, -1)
        torch.cuda.synchronize()
        assert torch.equal(got[0].view(torch.int16), got[1].view(torch.int16))


@pytest.mark.parametrize("n,H,W,Cin,Cout,unpadded", [
    (2, 13, 17, 64, 64, False), (1, 30, 50, 512, 512, False), (2, 9, 11, 128, 256, False), (1, 5, 7, 64, 128, True),
    (3, 38, 63, 256, 512, True), (1, 24, 40, 64, 128, False)])
def test_conv3x3_matches_fp32_convolution(dev, conv_variant, n, H, W, Cin, Cout, unpadded):
    """bf16 operands, fp32 accumulation: against torch's fp32 convolution of the SAME bf16-rounded operands the
    only differences are summation order and the bf16 rounding of the output (2^-9 relative)."""
    from aznet_b200 import ops
    g = torch.Generator().manual_seed(n * 1000 + H * 10 + Cin)
    x = torch.randn((n, H, W, Cin), generator=g).to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) * (2.0 / (9 * Cin)) ** 0.5).to(torch.bfloat16)
    b = torch.randn((Cout,), generator=g) * 0.1
    ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1)).permute(0, 2, 3, 1)
    xp = ops.nhwc_border(x.to(dev).contiguous(), True)
    wt = ops.pack_conv_weight(w.float().to(dev))
    shape = (n, H, W, Cout) if unpadded else (n, H + 2, W + 2, Cout)
    out = torch.full(shape, 5.0, dtype=torch.bfloat16, device=dev)
    y = ops.conv3x3(xp, wt, b.to(dev), relu=True, out=out, unpadded=unpadded)
    torch.cuda.synchronize()
    if not unpadded:
        yb = y.float().cpu()
        assert not yb[:, 0].any() and not yb[:, -1].any() and not yb[:, :, 0].any() and not yb[:, :, -1].any()
        y = ops.nhwc_border(y, False)
    got = y.float().cpu()
    err = (got - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2e-3
    assert bool((err <= tol).all()), (float(err.max()), float(ref.abs().max()))


@pytest.mark.parametrize("n,H,W,Cin,Cs,Cout,unpadded", [(2, 13, 17, 3, 8, 64, False), (1, 30, 50, 3, 8, 64, True), (2, 9, 11, 7, 8, 16, False),
                                                        (1, 24, 40, 1, 16, 128, False)])
def test_conv_patches_matches_fp32_convolution(dev, n, H, W, Cin, Cs, Cout, unpadded):
    """The first-layer route (azn_patches3x3 + azn_conv_patches_forward: the 3x3 neighbourhood gathered into one K = 64
    tap) against torch's fp32 convolution of the same bf16 operands, and bit-identical patch gathering."""
    from aznet_b200 import ops
    g = torch.Generator().manual_seed(n * 1000 + H * 10 + Cin)
    x = torch.randn((n, H, W, Cin), generator=g).to(torch.bfloat16)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) * (2.0 / (9 * Cin)) ** 0.5).to(torch.bfloat16)
    b = torch.randn((Cout,), generator=g) * 0.1
    ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1)).permute(0, 2, 3, 1)
    xs = torch.zeros((n, H + 2, W + 2, Cs), dtype=torch.bfloat16)
    xs[:, 1:H + 1, 1:W + 1, :Cin] = x
    patches = ops.patches3x3(xs.to(dev).contiguous(), Cin, 64)
    # the gather itself is a pure copy: compare with unfold on the host
    pad = torch.zeros((n, H + 4, W + 4, Cin), dtype=torch.bfloat16)
    pad[:, 2:H + 2, 2:W + 2] = x
    exp = torch.zeros((n, H + 2, W + 2, 64), dtype=torch.bfloat16)
    for ky in range(3):
        for kx in range(3):
            exp[:, 1:H + 1, 1:W + 1, (ky * 3 + kx) * Cin:(ky * 3 + kx + 1) * Cin] = pad[:, 1 + ky:1 + ky + H, 1 + kx:1 + kx + W]
    assert torch.equal(patches.cpu().view(torch.int16), exp.view(torch.int16))
    wt = ops.pack_patch_weight(w.float().to(dev), 64)
    shape = (n, H, W, Cout) if unpadded else (n, H + 2, W + 2, Cout)
    out = torch.full(shape, 5.0, dtype=torch.bfloat16, device=dev)
    y = ops.conv_patches(patches, wt, b.to(dev), relu=True, out=out, unpadded=unpadded)
    torch.cuda.synchronize()
    if not unpadded:
        yb = y.float().cpu()
        assert not yb[:, 0].any() and not yb[:, -1].any() and not yb[:, :, 0].any() and not yb[:, :, -1].any()
        y = ops.nhwc_border(y, False)
    got = y.float().cpu()
    err = (got - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2e-3
    assert bool((err <= tol).all()), (float(err.max()), float(ref.abs().max()))


@pytest.mark.parametrize("n,H,W,Cin,Cs,relu", [(2, 13, 17, 3, 8, True), (1, 30, 50, 3, 8, True), (3, 7, 5, 1, 8, False), (1, 480, 800, 3, 8, True),
                                                (2, 11, 9, 2, 8, True)])
def test_conv_direct_matches_fp32_convolution_and_the_patches_route(dev, n, H, W, Cin, Cs, relu):
    """conv1_1 in one kernel (azn_conv3x3_direct_forward: mma.sync over 16-pixel tiles, A fragments straight from the
    network input) against torch's fp32 convolution of the same bf16 operands -- the tolerance of the other conv tests --
    with a zero border, tiles that straddle rows / images / the end of the grid, and within
    one bf16 ulp of the patches route it replaces (same products, another summation order)."""
    from aznet_b200 import ops
    g = torch.Generator().manual_seed(n * 1000 + H * 10 + Cin)
    x = torch.randn((n, H, W, Cin), generator=g).to(torch.bfloat16)
    w = (torch.randn((64, Cin, 3, 3), generator=g) * (2.0 / (9 * Cin)) ** 0.5).to(torch.bfloat16)
    b = torch.randn((64,), generator=g) * 0.1
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    if relu:
        ref = torch.relu(ref)
    xs = torch.zeros((n, H + 2, W + 2, Cs), dtype=torch.bfloat16)
    xs[:, 1:H + 1, 1:W + 1, :Cin] = x
    xd = xs.to(dev).contiguous()
    wt = ops.pack_patch_weight(w.float().to(dev), 64)
    out = torch.full((n, H + 2, W + 2, 64), 5.0, dtype=torch.bfloat16, device=dev)
    y = ops.conv_direct(xd, Cin, wt, b.to(dev), relu=relu, out=out)
    torch.cuda.synchronize()
    yb = y.float().cpu()
    assert not yb[:, 0].any() and not yb[:, -1].any() and not yb[:, :, 0].any() and not yb[:, :, -1].any()
    got = yb[:, 1:H + 1, 1:W + 1]
    err = (got - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2e-3
    assert bool((err <= tol).all()), (float(err.max()), float(ref.abs().max()))
    if relu:
        old = ops.conv_patches(ops.patches3x3(xd, Cin, 64), wt, b.to(dev), relu=True).float().cpu()
        d = (old - yb).abs()
        assert bool((d <= 2.0 ** -7 * yb.abs() + 1e-3).all()), float(d.max())
    with pytest.raises(ValueError):
        ops.conv_direct(torch.zeros((1, 6, 6, 8), dtype=torch.bfloat16, device=dev), 4, wt, b.to(dev))        # 9 * 4 > 32
    with pytest.raises(ValueError):
        ops.conv_direct(torch.zeros((1, 6, 6, 16), dtype=torch.bfloat16, device=dev), 3, wt, b.to(dev))       # 16 channels per pixel


def test_conv3x3_argument_errors(dev):
    from aznet_b200 import ops
    x = torch.zeros((1, 6, 6, 32), dtype=torch.bfloat16, device=dev)
    wt = torch.zeros((64, 9 * 32), dtype=torch.bfloat16, device=dev)
    with pytest.raises(ValueError, match="multiple of 64"):
        ops.conv3x3(x, wt, torch.zeros(64, device=dev))


def test_vgg16_native_vs_oracle(dev, O):
    """The whole stack (13 convolutions, 4 pools) on two 96x128 images, full VGG16 widths, vs the fp32 CPU oracle
    on the same bf16-rounded weights and input.  Every layer rounds its output to bf16 (2^-9), so the error grows
    like sqrt(layers) * 2^-9 of the activation scale: asserted at 3 % of the map's maximum, 1 % on average."""
    from aznet_b200 import backbone, ops
    w = backbone.make_vgg16_weights(seed=5)
    wq = {k: (_bf(v[0]), v[1]) for k, v in w.items()}
    rng = np.random.RandomState(2)
    ims = rng.randint(0, 256, (2, 96, 128, 3)).astype(np.uint8)
    bb = backbone.VGG16Native(w, dev)
    _, blob = ops.image_blob(torch.from_numpy(ims).to(dev), 1.0, backbone.PIXEL_MEANS, 64, want_f32=True)
    got = bb.from_images(torch.from_numpy(ims).to(dev), 1.0).float().cpu().numpy()            # [2, 6, 8, 512]
    ref = O.vgg16_conv5(wq, _bf(blob.cpu().numpy())).transpose(0, 2, 3, 1)
    assert got.shape == ref.shape == (2, 6, 8, 512)
    scale = np.abs(ref).max()
    err = np.abs(got - ref)
    assert err.max() <= 0.03 * scale and err.mean() <= 0.01 * np.abs(ref).mean() + 1e-6, (err.max(), scale, err.mean())
    # the blob-level protocol of Net (f32 NCHW in and out) gives the same map
    via_data = bb(blob).cpu().numpy().transpose(0, 2, 3, 1)
    assert np.array_equal(via_data, got)
    # and the cuDNN library path agrees to the same tolerance (information, not the oracle)
    lib = backbone.VGG16Torch(w, dev)(blob).cpu().numpy().transpose(0, 2, 3, 1)
    assert np.abs(lib - ref).max() <= 0.03 * scale


def test_reduced_width_network_pads_channels(dev, O):
    """Channel counts below 64 are zero-padded inside the packed weights; the result must not change."""
    from aznet_b200 import backbone
    w = backbone.make_vgg16_weights(seed=5, width_div=8)
    wq = {k: (_bf(v[0]), v[1]) for k, v in w.items()}
    x = np.random.RandomState(4).standard_normal((1, 3, 75, 125)).astype(np.float32) * 50
    bb = backbone.VGG16Native(w, dev)
    got = bb(torch.from_numpy(x).to(dev)).cpu().numpy()
    ref = O.vgg16_conv5(wq, _bf(x))
    assert got.shape == ref.shape == (1, 64, 5, 8)
    assert np.abs(got - ref).max() <= 0.03 * np.abs(ref).max()
