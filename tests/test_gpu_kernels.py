"""GPU suite (run on the B200 box: pytest -m gpu): kernel-level parity of every C-ABI entry point
against the oracle on the same seeded inputs.  Bit-exact for ROI pooling, NMS and region
arithmetic; stated tolerances for the bf16 tensor-core layers and the f32-exp box decode."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from aznet_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu


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


def _natural_regions(O, n_levels=5):
    lv = np.array([[0, 0, 999, 599.]])
    out = [lv]
    for _ in range(n_levels - 1):
        lv = O.divide_region(lv, 10.0)
        out.append(lv)
    return np.vstack(out)


def _edge_rois(n_img):
    r = synth.make_rois(400, 600, 1000, seed=5, n_img=n_img)
    r[:20, 1:] = np.round(r[:20, 1:] / 8) * 8        # x*1/16 lands on .5: round half away from zero
    r[20:26, 1:] += 2000                             # outside the map -> empty bins -> 0
    r[26] = [0, 300, 300, 100, 100]                  # malformed -> 1x1
    r[27] = [0, 0, 0, 999, 599]                      # whole image
    r[28] = [0, 0, 0, 0, 0]
    r[29] = [0, -50, -50, 30, 30]                    # negative start: clamped
    return r


# ------------------------------------------------------------------------------- ROI pooling
@pytest.fixture(params=[0, 1, 501, 22, 422, 322, 42], ids=["auto", "direct", "direct_per_roi", "staged", "staged_pairs", "staged_bands", "staged_grouped"])
def pool_mode(request):
    """Kernel choice of azn_roi_pool_fwd: automatic, direct (L2-fed) kernels only, the direct kernel with one CTA per ROI (501), shared-memory-staged kernel with the
    per-ROI loop, the same over a two-level map (slice + row-pair table where both fit: an A/B variant), the same with row bands of 128-byte slices when the whole map does not fit (the 38x63 maps below: an
    A/B variant), staged kernel with the ROIs grouped by width whenever that path applies."""
    from aznet_b200 import _lib
    _lib.lib().azn_roi_pool_tune(request.param)
    yield request.param
    _lib.lib().azn_roi_pool_tune(20)


@pytest.mark.parametrize("hw", [(38, 63), (30, 50)])
def test_roi_pool_nchw_f32_bit_exact(dev, O, hw, pool_mode):
    from aznet_b200 import ops
    feat = synth.make_conv_maps(2, 64, hw[0], hw[1], seed=7)
    feat[1] -= 0.3                                   # negative values too: the layer itself is sign-agnostic
    rois = _edge_rois(2)
    ref, ref_am = O.roi_pool_fwd(feat, rois, want_argmax=True)
    got, am = ops.roi_pool(torch.from_numpy(feat).to(dev), torch.from_numpy(rois).to(dev), layout="NCHW", want_argmax=True)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(am.cpu().numpy(), ref_am)
    got2 = ops.roi_pool(torch.from_numpy(feat).to(dev), torch.from_numpy(rois).to(dev), layout="NCHW")     # no argmax
    assert np.array_equal(got2.cpu().numpy().view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("C", [64, 512])
def test_roi_pool_nhwc_bit_exact_f32_and_bf16(dev, O, C, pool_mode):
    from aznet_b200 import ops
    n_img = 2 if C == 512 else 3
    feat = synth.make_conv_maps(n_img, C, 38, 63, seed=9)
    rois = _edge_rois(n_img)
    if C == 512:
        nat = _natural_regions(O)                    # the 739 natural full-zoom regions of a 600x1000 image
        rois = np.vstack([rois, np.hstack([np.zeros((len(nat), 1)), nat]).astype(np.float32)])
    ref = O.roi_pool_fwd(feat, rois).transpose(0, 2, 3, 1)          # -> [R, ph, pw, C]
    f = torch.from_numpy(feat).to(dev)
    r = torch.from_numpy(rois).to(dev)
    got = ops.roi_pool(f.permute(0, 2, 3, 1).contiguous(), r, layout="NHWC")
    assert np.array_equal(got.cpu().numpy().view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))
    # bf16: pool(bf16(x)) == bf16(pool(x)) bit for bit
    fb = f.to(torch.bfloat16)
    ref_b = O.roi_pool_fwd(fb.float().cpu().numpy(), rois).transpose(0, 2, 3, 1)
    got_b = ops.roi_pool(fb.permute(0, 2, 3, 1).contiguous(), r, layout="NHWC")
    assert np.array_equal(got_b.float().cpu().numpy(), ref_b)
    # the conversion kernel produces the same NHWC bf16 map
    conv = ops.nchw_to_nhwc_bf16(f)
    assert torch.equal(conv, fb.permute(0, 2, 3, 1).contiguous())
    # bf16 NCHW variant
    got_c = ops.roi_pool(fb.contiguous(), r, layout="NCHW")
    assert np.array_equal(got_c.float().cpu().numpy(), ref_b.transpose(0, 3, 1, 2))


def test_roi_pool_device_count_bad_index_and_empty(dev, O, pool_mode):
    from aznet_b200 import ops
    feat = synth.make_conv_maps(1, 64, 20, 20, seed=1)
    rois = synth.make_rois(50, 300, 300, seed=2)
    rois[3, 0] = 7                                   # batch index out of range -> zero row (documented)
    f = torch.from_numpy(feat).to(dev).permute(0, 2, 3, 1).contiguous()
    r = torch.from_numpy(rois).to(dev)
    out = torch.full((50, 7, 7, 64), -1.0, device=dev)
    ops.roi_pool(f, r, layout="NHWC", n_rois=torch.tensor([10], dtype=torch.int32, device=dev), out=out)
    got = out.cpu().numpy()
    ok = rois[:10].copy()
    ok[3, 0] = 0
    ref = O.roi_pool_fwd(feat, ok).transpose(0, 2, 3, 1)
    ref[3] = 0
    assert np.array_equal(got[:10], ref) and np.all(got[10:] == -1.0)   # rows past the live count untouched
    z = ops.roi_pool(f, r[:0].contiguous(), layout="NHWC")
    assert z.shape[0] == 0
    with pytest.raises(ValueError):
        ops.roi_pool(torch.zeros((1, 4, 4, 3), device=dev), r, layout="NHWC")     # C*4 % 16 != 0


def test_roi_pool_per_roi_kernel_choice_same_bits(dev, O):
    """azn_roi_pool_fwd_ex(kernel_choice = 3): the direct kernel with one CTA per ROI (what the search engine asks for on its
    deep levels: many small ROIs, device-side count) -- same bits as the default direct kernel and as the oracle, f32 and bf16,
    rows past the live count untouched, edge ROIs and a bad batch index included."""
    from aznet_b200 import ops
    n_img, C = 3, 64
    feat = synth.make_conv_maps(n_img, C, 30, 50, seed=11)
    feat[1] -= 0.25
    rois = np.vstack([synth.make_rois(700, 600, 1000, seed=9, n_img=n_img), _edge_rois(n_img)]).astype(np.float32)
    rois[:, 1:] *= 0.8                                 # the voc.yml scale: 30 x 50 maps
    rois[5, 0] = n_img + 1                             # bad batch index -> zero row
    ok = rois.copy()
    ok[5, 0] = 0
    live = rois.shape[0] - 13
    ref = O.roi_pool_fwd(feat, ok).transpose(0, 2, 3, 1).copy()
    ref[5] = 0
    f = torch.from_numpy(feat).to(dev).permute(0, 2, 3, 1).contiguous()
    r = torch.from_numpy(rois).to(dev)
    n = torch.tensor([live], dtype=torch.int32, device=dev)
    for fm in (f, f.to(torch.bfloat16)):
        outs = []
        for per_roi in (False, True):
            out = torch.full((rois.shape[0], 7, 7, C), 7.0, dtype=fm.dtype, device=dev)
            ops.roi_pool(fm, r, layout="NHWC", n_rois=n, out=out, per_roi=per_roi)
            outs.append(out.float().cpu().numpy())
        assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
        assert np.all(outs[1][live:] == 7.0)
        if fm.dtype == torch.float32:
            assert np.array_equal(outs[1][:live].view(np.uint32), np.ascontiguousarray(ref[:live]).view(np.uint32))


@pytest.mark.parametrize("n_img", [1, 3])
def test_roi_pool_geometry_prepass_large_R(dev, O, n_img):
    """The geometry pre-pass variant of the staged NHWC kernel (azn_roi_pool_tune(922), from 4096 ROIs: the ROIs' bin bounds come
    from the records of roi_geom_kernel instead of being derived once per channel slice; an A/B variant with an instantiation of
    its own): same device functions, so the same bits -- against the oracle, and against the default kernel without the
    pre-pass; edge ROIs and bad batch indices included."""
    from aznet_b200 import _lib, ops
    C = 64
    feat = synth.make_conv_maps(n_img, C, 38, 63, seed=21)
    feat[0] -= 0.25
    rois = np.vstack([synth.make_rois(5000, 600, 1000, seed=5, n_img=n_img), _edge_rois(n_img)]).astype(np.float32)
    rois[17, 0] = n_img + 3                           # bad batch index -> zero row
    ok = rois.copy()
    ok[17, 0] = 0
    ref = O.roi_pool_fwd(feat, ok).transpose(0, 2, 3, 1).copy()
    ref[17] = 0
    f = torch.from_numpy(feat).to(dev).permute(0, 2, 3, 1).contiguous()
    r = torch.from_numpy(rois).to(dev)
    lib = _lib.lib()
    try:
        outs = {}
        for mode in (22, 922):
            lib.azn_roi_pool_tune(mode)
            outs[mode] = ops.roi_pool(f, r, layout="NHWC").cpu().numpy()
            outs[(mode, "bf16")] = ops.roi_pool(f.to(torch.bfloat16), r, layout="NHWC").float().cpu().numpy()
    finally:
        lib.azn_roi_pool_tune(20)
    assert np.array_equal(outs[22].view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))
    assert np.array_equal(outs[22].view(np.uint32), outs[922].view(np.uint32))
    assert np.array_equal(outs[922].view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))
    assert np.array_equal(outs[(22, "bf16")].view(np.uint32), outs[(922, "bf16")].view(np.uint32))


@pytest.mark.parametrize("n_img,C,hw", [(1, 512, (38, 63)), (3, 64, (38, 63)), (2, 64, (75, 40)), (64, 64, (38, 63))])
def test_roi_pool_row_bands_equal_the_direct_kernel(dev, O, n_img, C, hw):
    """The banded staged path (azn_roi_pool_tune(322); a pre-pass sorts the ROIs by the row band that holds them, 128-byte slices per band, a second
    launch for the ROIs no band holds) against the direct kernel -- itself pinned to the oracle above -- on thousands of
    ROIs: every band, ROIs taller than a band, outside the map, bad batch indices, a device-side count, a map with -0 and
    NaN (the exact compare-select fallback), f32 and bf16; a sample against the oracle as well."""
    from aznet_b200 import _lib, ops
    H, W = hw
    R = 6000 if n_img < 64 else 64 * 80
    feat = synth.make_conv_maps(n_img, C, H, W, seed=31)
    rois = synth.make_rois(R, H * 16, W * 16, seed=17, n_img=n_img)
    rois[::211, 0] = n_img + 2                                    # bad batch index -> zero row
    rois[7:12, 1:] += 5000                                        # outside the map
    rois[12] = [0, 0, 0, W * 16 - 1, H * 16 - 1]                  # the whole map: taller than any band
    rois[13] = [0, 40, 0, 90, H * 16 - 1]                         # a full-height sliver
    f = torch.from_numpy(feat).to(dev).permute(0, 2, 3, 1).contiguous()
    r = torch.from_numpy(rois).to(dev)
    lib = _lib.lib()
    try:
        for variant in ("relu", "signed"):
            if variant == "signed":
                g = feat.copy()
                g[g < 0.4] = 0.0
                rs = np.random.RandomState(5)
                g[rs.rand(*g.shape) < 0.2] *= -1.0                # -0.0 and negative values
                g[rs.rand(*g.shape) < 0.01] = np.nan
                f = torch.from_numpy(g).to(dev).permute(0, 2, 3, 1).contiguous()
            for dt in (torch.float32, torch.bfloat16):
                fm = f.to(dt)
                lib.azn_roi_pool_tune(1)
                want = ops.roi_pool(fm, r, layout="NHWC")
                lib.azn_roi_pool_tune(322)
                got = ops.roi_pool(fm, r, layout="NHWC")
                iv = torch.int32 if dt == torch.float32 else torch.int16
                assert torch.equal(got.view(iv), want.view(iv)), (variant, dt)
                # a device-side count (the staged kernel asked for explicitly): rows past it stay untouched
                out = torch.full_like(want, -1.0)
                live = R // 3
                ops.roi_pool(fm, r, layout="NHWC", n_rois=torch.tensor([live], dtype=torch.int32, device=dev), out=out, staged=True)
                assert torch.equal(out[:live].view(iv), want[:live].view(iv)) and bool((out[live:] == -1.0).all())
        # and the oracle on a sample (f32, the ReLU-like map)
        idx = np.r_[0:40, R - 40:R]
        ok = rois[idx].copy()
        bad = (ok[:, 0] < 0) | (ok[:, 0] >= n_img)
        ok[bad, 0] = 0
        ref = O.roi_pool_fwd(feat, ok).transpose(0, 2, 3, 1).copy()
        ref[bad] = 0
        lib.azn_roi_pool_tune(322)
        got = ops.roi_pool(torch.from_numpy(feat).to(dev).permute(0, 2, 3, 1).contiguous(), r, layout="NHWC")
        assert np.array_equal(got[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint32), ref.view(np.uint32))
    finally:
        lib.azn_roi_pool_tune(20)


def test_roi_pool_signed_zero_nan_and_many_images(dev, O, pool_mode):
    """`v > best ? v : best` semantics: the first of +-0 ties wins, NaN never wins; ROIs of 5 images interleaved,
    bad batch indices in the middle of the list (they go to their own bucket of the staged kernel)."""
    from aznet_b200 import ops
    n_img = 5
    feat = synth.make_conv_maps(n_img, 32, 30, 50, seed=21)
    feat[feat < 0.4] = 0.0
    rs = np.random.RandomState(5)
    feat[rs.rand(*feat.shape) < 0.2] *= -1.0                      # -0.0 and negative values
    feat[rs.rand(*feat.shape) < 0.01] = np.nan
    rois = synth.make_rois(700, 480, 800, seed=8, n_img=n_img)
    rois[::97, 0] = n_img + 3
    rois[5, 0] = -1
    ok = rois.copy()
    bad = (ok[:, 0] < 0) | (ok[:, 0] >= n_img)
    ok[bad, 0] = 0
    ref = O.roi_pool_fwd(feat, ok)
    ref[bad] = 0
    f = torch.from_numpy(feat).to(dev)
    r = torch.from_numpy(rois).to(dev)
    got = ops.roi_pool(f, r, layout="NCHW").cpu().numpy()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    got = ops.roi_pool(f.permute(0, 2, 3, 1).contiguous(), r, layout="NHWC").cpu().numpy()
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref.transpose(0, 2, 3, 1)).view(np.uint32))
    fb = f.to(torch.bfloat16)
    ref_b = O.roi_pool_fwd(fb.float().cpu().numpy(), ok)
    ref_b[bad] = 0
    ref_b = torch.from_numpy(ref_b).to(torch.bfloat16).float().numpy()       # an all-NaN window leaves -FLT_MAX -> bf16 -inf
    got = ops.roi_pool(fb.permute(0, 2, 3, 1).contiguous(), r, layout="NHWC").float().cpu().numpy()
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref_b.transpose(0, 2, 3, 1)).view(np.uint32))


# ------------------------------------------------------------------------------- NMS
def test_nms_golden_and_oracle(dev, O, golden):
    from aznet_b200 import ops
    g = golden["nms"]
    for n in (1, 2, 17, 300, 2000):
        d = synth.make_dets(n, seed=3)
        for th in (0.3, 0.5, 0.7):
            keep, cnt = ops.nms(torch.from_numpy(d).to(dev), th)
            got = keep[:int(cnt.item())].cpu().tolist()
            assert got == list(g["keep_n%d_t%d" % (n, int(th * 10))]), (n, th)
    for th, key in ((0.5, "edge_keep_t5"), (0.51, "edge_keep_t51")):
        keep, cnt = ops.nms(torch.from_numpy(g["edge_dets"]).to(dev), th)
        assert keep[:int(cnt.item())].cpu().tolist() == list(g[key])
    keep, cnt = ops.nms(torch.zeros((0, 5), device=dev), 0.5)
    assert int(cnt.item()) == 0


@pytest.mark.parametrize("n,th", [(8000, 0.3), (8000, 0.7), (20000, 0.3), (20000, 0.7), (4096, 0.5), (2049, 0.5), (3000, 0.4), (21000, 0.5), (33000, 0.6)])
def test_nms_large_matches_oracle(dev, O, n, th):
    from aznet_b200 import ops
    d = synth.make_dets(n, seed=3)
    keep, cnt = ops.nms(torch.from_numpy(d).to(dev), th)
    assert keep[:int(cnt.item())].cpu().tolist() == O.nms(d, th)


@pytest.mark.parametrize("n", [1, 31, 33, 255, 2047, 2048])
def test_nms_small_sort_with_ties(dev, O, n):
    """n <= 2048 sorts with eight threads per detection (nms_rank1_kernel): sizes around its group / CTA / tile borders,
    many exact score ties and +-0, against the oracle and against the round-1 one-thread-per-detection kernel."""
    from aznet_b200 import _lib, ops
    d = synth.make_dets(n, seed=23 + n)
    d[:, 4] = np.round(d[:, 4] * 16) / 16 - 0.25
    d[::5, 4] = 0.0
    d[2::5, 4] = -0.0
    d = np.ascontiguousarray(d, dtype=np.float32)
    ref = O.nms(d, 0.45)
    dt = torch.from_numpy(d).to(dev)
    lib = _lib.lib()
    try:
        for mode in (0, 8):
            lib.azn_nms_tune(mode)
            keep, cnt = ops.nms(dt, 0.45)
            assert keep[:int(cnt.item())].cpu().tolist() == ref, (n, mode)
    finally:
        lib.azn_nms_tune(0)


def test_nms_ties_and_idempotence(dev, O):
    from aznet_b200 import ops
    d = synth.make_dets(600, seed=11)
    d[:, 4] = np.round(d[:, 4] * 20) / 20            # many exact score ties: order = stable argsort reversed
    keep, cnt = ops.nms(torch.from_numpy(d).to(dev), 0.5)
    k1 = keep[:int(cnt.item())].cpu().tolist()
    assert k1 == O.nms(d, 0.5)
    kept = np.ascontiguousarray(d[k1])
    keep2, cnt2 = ops.nms(torch.from_numpy(kept).to(dev), 0.5)     # survivors never suppress each other
    assert int(cnt2.item()) == len(k1)


@pytest.mark.parametrize("dist", ["ties", "all_equal", "saturated", "negative_and_zero", "wide_range"])
def test_nms_bucket_sort_orders_like_the_rank_sort(dev, O, dist):
    """n >= 4096 sorts by score buckets (nms_bucket_kernel + nms_bucket_rank_kernel) instead of the all-pairs rank: the keep
    list equals the oracle's and the all-pairs sort's (azn_nms_tune(8)) on score distributions that stress the bucketing --
    thousands of exact ties, one single score, scores saturated at 1.0, negative scores with -0 / +0, twenty decades."""
    from aznet_b200 import _lib, ops
    n = 6000
    d = synth.make_dets(n, seed=13)
    rng = np.random.default_rng(3)
    if dist == "ties":
        d[:, 4] = np.round(d[:, 4] * 40) / 40
    elif dist == "all_equal":
        d[:, 4] = 0.5
    elif dist == "saturated":
        d[:, 4] = np.where(rng.random(n) < 0.6, 1.0, d[:, 4]).astype(np.float32)
    elif dist == "negative_and_zero":
        d[:, 4] = (d[:, 4] - 0.5).astype(np.float32)
        d[::7, 4] = 0.0
        d[3::7, 4] = -0.0
    else:
        d[:, 4] = np.exp(rng.uniform(-40, 5, n)).astype(np.float32)
    d = np.ascontiguousarray(d, dtype=np.float32)
    dt = torch.from_numpy(d).to(dev)
    ref = O.nms(d, 0.5)
    keep, cnt = ops.nms(dt, 0.5)
    assert keep[:int(cnt.item())].cpu().tolist() == ref, dist
    lib = _lib.lib()
    try:
        for mode in (8, 16, 24, 32, 56, 64):           # all-pairs sort / tile-by-tile greedy pass / both / float32 mask kernel / all / persistent greedy chain: the A/B variants
            lib.azn_nms_tune(mode)
            keep2, cnt2 = ops.nms(dt, 0.5)
            assert keep2[:int(cnt2.item())].cpu().tolist() == ref, (dist, mode)
    finally:
        lib.azn_nms_tune(0)


def test_nms_long_dependency_chain(dev, O):
    """The block-wise greedy pass decides a super-tile of 1024 rows in rounds; a staircase of boxes in which every box is
    suppressed by its predecessor alone (kept, removed, kept, ...) is its worst case -- one round per row -- and must
    still give the serial scan's answer.  Two staircases across super-tile borders plus unrelated boxes in between."""
    from aznet_b200 import ops
    n = 3000
    d = np.zeros((n, 5), np.float32)
    x = np.arange(n, dtype=np.float32) * 4.0                      # 10-wide boxes shifted by 4: IoU(i, i+1) = 7/15 > 0.4, IoU(i, i+2) < 0.4
    d[:, 0], d[:, 2] = x, x + 10.0
    d[:, 1], d[:, 3] = 5.0, 50.0
    d[:, 4] = np.linspace(1.0, 0.1, n)                            # score order = position
    d[1000:1400, 1] += 500.0                                      # an unrelated group on another row of the image
    d[1000:1400, 3] += 500.0
    d = np.ascontiguousarray(d)
    ref = O.nms(d, 0.4)
    assert 1200 < len(ref) < 1800
    keep, cnt = ops.nms(torch.from_numpy(d).to(dev), 0.4)
    assert keep[:int(cnt.item())].cpu().tolist() == ref


def test_nms_inside_cuda_graph_and_on_side_stream(dev, O):
    """azn_nms forks its greedy chain onto an internal stream; inside a stream capture it stays on the caller's stream.
    Both give the oracle's keep list, also when replayed from a graph and when the caller's stream is not the default."""
    from aznet_b200 import ops
    d = synth.make_dets(5000, seed=9)
    ref = O.nms(d, 0.5)
    dt = torch.from_numpy(d).to(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        keep, cnt = ops.nms(dt, 0.5)                               # warm-up on the side stream (two-stream path)
        side.synchronize()
        assert keep[:int(cnt.item())].cpu().tolist() == ref
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            keep_g, cnt_g = ops.nms(dt, 0.5)
    for _ in range(2):
        keep_g.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert keep_g[:int(cnt_g.item())].cpu().tolist() == ref


def test_nms_batched_matches_per_segment(dev, O):
    from aznet_b200 import ops
    rng = np.random.default_rng(0)
    sizes = [0, 1, 100, 37, 64, 65, 100, 250, 3, 1000] + [int(x) for x in rng.integers(0, 101, 150)]
    segs = [synth.make_dets(max(s, 1), seed=50 + i)[:s] for i, s in enumerate(sizes)]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    d = np.ascontiguousarray(np.vstack(segs).astype(np.float32))
    keep, cnt = ops.nms_batched(torch.from_numpy(d).to(dev), torch.from_numpy(off).to(dev), 0.5)
    keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
    for i, s in enumerate(sizes):
        ref = O.nms(np.ascontiguousarray(segs[i]), 0.5) if s else []
        assert list(keep[off[i]:off[i] + cnt[i]]) == ref, i


# ------------------------------------------------------------------------------- regions / decode
def test_divide_region_bit_exact(dev, O, golden):
    from aznet_b200 import ops
    g = golden["div"]
    for k in g.files:
        if not k.startswith("in_"):
            continue
        out, cnt = ops.divide_region(torch.from_numpy(np.ascontiguousarray(g[k], dtype=np.float64)).to(dev), 10.0)
        got = out[:int(cnt.item())].cpu().numpy()
        assert np.array_equal(got.view(np.uint64), g["out_" + k[3:]].view(np.uint64)), k
    lv = torch.from_numpy(g["in_root_600x1000"]).to(dev)
    for lvl in range(2, 6):
        out, cnt = ops.divide_region(lv, 10.0)
        lv = out[:int(cnt.item())].contiguous()
        assert np.array_equal(lv.cpu().numpy(), g["cascade_%d" % lvl])
    out, cnt = ops.divide_region(torch.from_numpy(g["sift_in"]).to(dev), 10.0, sift_only=True)
    assert np.array_equal(out[:int(cnt.item())].cpu().numpy(), g["sift_out"])
    out, cnt = ops.divide_region(torch.zeros((0, 4), dtype=torch.float64, device=dev), 10.0)
    assert int(cnt.item()) == 0


def test_decode_boxes_within_1e5(dev, O, golden):
    from aznet_b200 import ops
    g = golden["search"]
    got = ops.decode_boxes(torch.from_numpy(g["bbox_boxes"]).to(dev), torch.from_numpy(g["bbox_deltas"]).to(dev), 600, 1000)
    ref = g["bbox_clip"]
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-5, atol=1e-3)      # north_star: 1e-5 relative (to the box scale)


# ------------------------------------------------------------------------------- fc layers (tcgen05)
def _fc_case(dev, M, N, K, act, seed, m_live=None, out_dtype=torch.bfloat16, act_aux=0):
    from aznet_b200 import ops
    g = torch.Generator().manual_seed(seed)
    A = (torch.randn((M, K), generator=g) * 0.5).to(torch.bfloat16)
    W = (torch.randn((N, K), generator=g) * (1.0 / np.sqrt(K))).to(torch.bfloat16)
    b = torch.randn((N,), generator=g) * 0.1
    ml = None if m_live is None else torch.tensor([m_live], dtype=torch.int32, device=dev)
    out = torch.full((M, (N + 7) // 8 * 8), -7.0, dtype=out_dtype, device=dev)[:, :N] if out_dtype == torch.float32 else None
    if out is not None:
        got = ops.fc_forward(A.to(dev), W.to(dev), b.to(dev), act, act_aux, out_dtype=out_dtype, m_live=ml, out=out)
    else:
        got = ops.fc_forward(A.to(dev), W.to(dev), b.to(dev), act, act_aux, out_dtype=out_dtype, m_live=ml)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + b                       # plain fp32 reference of the same op
    return got.float().cpu(), ref


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 256, 128), (64, 256, 3136), (300, 512, 1024),
                                   (1000, 4096, 3136), (2500, 1280, 4096), (700, 56, 1280), (129, 105, 256)])
def test_fc_forward_matches_fp32_reference(dev, M, N, K):
    from aznet_b200 import _lib as L
    got, ref = _fc_case(dev, M, N, K, L.ACT_RELU, seed=M + N + K)
    ref = torch.relu(ref)
    err = (got - ref).abs().max().item()
    # bf16 output rounding (2^-9 relative) on |values| <~ 4, fp32 accumulation: stated tolerance 3e-2 abs
    assert err < 3e-2, "max abs err %.4g" % err
    assert (got - ref).abs().mean().item() < 3e-3


def test_fc_forward_live_count_and_f32_heads(dev):
    from aznet_b200 import _lib as L
    nsub = 11
    got, ref = _fc_case(dev, 900, 56, 1280, L.ACT_AZ_HEAD, seed=5, m_live=333, out_dtype=torch.float32, act_aux=nsub)
    sig = torch.sigmoid(ref)
    want = ref.clone()
    want[:, :nsub] = sig[:, :nsub]
    want[:, 5 * nsub] = sig[:, 5 * nsub]
    assert (got[:333] - want[:333]).abs().max().item() < 2e-3          # f32 out of bf16 operands
    assert torch.all(got[333:] == -7.0)                                # rows past the live count untouched


@pytest.mark.parametrize("M,m_live", [(900, 333), (64, None), (1, None), (1536, 1494), (130, 65)])
def test_az_heads_kernel_vs_fp32_and_the_gemm_epilogue(dev, M, m_live):
    """azn_az_heads_forward (mma.sync kernel) = fp32 product of the same bf16 operands + bias + sigmoids on columns
    [0, nsub) and 5*nsub, to 2e-3 (fp32 accumulation of bf16 products; only the summation order differs), rows past the
    live count untouched; and it agrees with azn_fc_forward's AZN_ACT_AZ_HEAD epilogue to 1e-4."""
    from aznet_b200 import _lib as L
    from aznet_b200 import ops
    nsub, K = 11, 1280
    N = 5 * nsub + 1
    gen = torch.Generator(device="cpu").manual_seed(7 + M)
    A = torch.relu(torch.randn((M, K), generator=gen)).to(torch.bfloat16).to(dev)
    W = (torch.randn((N, K), generator=gen) * 0.05).to(torch.bfloat16).to(dev)
    b = (torch.randn((N,), generator=gen) * 0.1).to(dev)
    ml = None if m_live is None else torch.tensor([m_live], dtype=torch.int32, device=dev)
    live = M if m_live is None else m_live
    out = torch.full((M, 56), -7.0, dtype=torch.float32, device=dev)
    ops.az_heads(A, W, b, nsub, m_live=ml, out=out)
    ref = A.float() @ W.float().t() + b
    want = ref.clone()
    want[:, :nsub] = torch.sigmoid(ref[:, :nsub])
    want[:, 5 * nsub] = torch.sigmoid(ref[:, 5 * nsub])
    assert (out[:live] - want[:live]).abs().max().item() < 2e-3
    assert torch.all(out[live:] == -7.0)
    out2 = torch.full((M, 56), -7.0, dtype=torch.float32, device=dev)
    ops.fc_forward(A, W, b, L.ACT_AZ_HEAD, nsub, m_live=ml, out=out2)
    assert (out[:live] - out2[:live]).abs().max().item() < 1e-4


def test_fc_forward_splitk_small_m(dev):
    from aznet_b200 import _lib as L
    for M in (1, 8, 64, 200):
        got, ref = _fc_case(dev, M, 4096, 25088, L.ACT_RELU, seed=M)
        assert (got - torch.relu(ref)).abs().max().item() < 4e-2, M


def test_fc_forward_softmax_rows(dev):
    from aznet_b200 import _lib as L
    C = 21
    got, ref = _fc_case(dev, 300, 5 * C, 4096, L.ACT_SOFTMAX_BBOX, seed=9, out_dtype=torch.float32, act_aux=C)
    want = ref.clone()
    want[:, :C] = torch.softmax(ref[:, :C], dim=1)
    assert (got - want).abs().max().item() < 5e-3
    np.testing.assert_allclose(got[:, :C].sum(1).numpy(), 1.0, atol=1e-5)


def test_fc_forward_argument_errors(dev):
    from aznet_b200 import ops
    A = torch.zeros((128, 100), dtype=torch.bfloat16, device=dev)
    W = torch.zeros((64, 100), dtype=torch.bfloat16, device=dev)
    with pytest.raises(ValueError):
        ops.fc_forward(A, W, torch.zeros(64, device=dev))              # K % 64 != 0
