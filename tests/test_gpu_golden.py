"""GPU suite: the CUDA path against fixtures produced by the REFERENCE'S OWN CODE (tests/golden/, made by
oracle/gen_golden.py): Forward_cpu of the unmodified roi_pooling_layer.cpp / pooling_layer.cpp, and the
reference's im_detect + test_net + apply_nms; plus the headline configuration at full width (512-channel maps,
25088 -> 4096 -> 1280 -> 56 heads, the bench's weights) against the oracle with the tolerance stated in the test."""
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


def _layer_feat():
    feat = synth.make_conv_maps(2, 8, 38, 63, seed=7)          # as oracle/gen_golden.py::gen_layers
    feat[1] -= 0.5
    feat[1, 0, 3, 4] = np.nan
    feat[1, 1, 10:20, 10:30] = -0.0
    return feat


@pytest.mark.parametrize("mode", [0, 1, 501, 22, 322, 42], ids=["auto", "direct", "direct_per_roi", "staged", "staged_bands", "staged_grouped"])
def test_roi_pool_equals_reference_layer_golden(dev, golden, mode):
    """azn_roi_pool_fwd == ROIPoolingLayer<float>::Forward_cpu of the reference's own source, bit for bit: values
    and argmax (NCHW f32, the layer's blob layout), NHWC f32, and bf16 (max commutes with the monotone rounding)."""
    from aznet_b200 import _lib, ops
    g = golden["caffe_layers"]
    feat = _layer_feat()
    f = torch.from_numpy(feat).to(dev)
    _lib.lib().azn_roi_pool_tune(mode)
    try:
        for tag in ("edge", "natural"):
            rois = torch.from_numpy(g["roi_%s_rois" % tag]).to(dev)
            ref, ref_am = g["roi_%s_out" % tag], g["roi_%s_argmax" % tag].astype(np.int32)
            got, am = ops.roi_pool(f, rois, layout="NCHW", want_argmax=True)
            assert np.array_equal(got.cpu().numpy().view(np.uint32), ref.view(np.uint32)), tag
            assert np.array_equal(am.cpu().numpy(), ref_am), tag
            got = ops.roi_pool(f.permute(0, 2, 3, 1).contiguous(), rois, layout="NHWC")
            assert np.array_equal(got.permute(0, 3, 1, 2).cpu().numpy().view(np.uint32), ref.view(np.uint32)), tag
            # bf16 storage: the reference layer applied to the bf16-rounded map, rounded -- equal to rounding its output
            # wherever the map has no NaN (NaN never wins in either)
            fb = f.to(torch.bfloat16)
            got_b = ops.roi_pool(fb.permute(0, 2, 3, 1).contiguous(), rois, layout="NHWC").permute(0, 3, 1, 2).float().cpu().numpy()
            want_b = torch.from_numpy(ref).to(torch.bfloat16).float().numpy()
            assert np.array_equal(got_b.view(np.uint32), want_b.view(np.uint32)), tag
    finally:
        _lib.lib().azn_roi_pool_tune(20)


def test_roi_pool_microbench_size_vs_oracle_rows(dev, O):
    """BASELINE config #4 at full size -- 20 000 ROIs of the microbench's size mix over a 512 x 38 x 63 map, the staged
    kernel with its geometry pre-pass, 1 GB of output -- against the oracle on 600 of the rows (a ROI's pooled row depends
    on that ROI alone, so a sample of rows is compared bit for bit: bf16 and f32), and two size-independent properties
    over ALL rows: every pooled value is one of the map's values of its channel or +0 (an empty bin), and pooling the
    ROIs in reversed order gives the reversed output."""
    from aznet_b200 import ops
    C, H, W, R = 512, 38, 63, 20000
    feat = synth.make_conv_maps(1, C, H, W, seed=7)
    rois = synth.make_rois(R, 600, 1000, seed=3)
    rng = np.random.RandomState(11)
    pick = np.sort(rng.choice(R, 600, replace=False))
    f = torch.from_numpy(feat).to(dev).permute(0, 2, 3, 1).contiguous()
    r = torch.from_numpy(rois).to(dev)
    ref = O.roi_pool_fwd(feat, rois[pick]).transpose(0, 2, 3, 1)
    got = ops.roi_pool(f, r, layout="NHWC")
    assert np.array_equal(got[torch.from_numpy(pick).to(dev)].cpu().numpy().view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))
    fb = f.to(torch.bfloat16)
    got_b = ops.roi_pool(fb, r, layout="NHWC")
    ref_b = O.roi_pool_fwd(fb.float().permute(0, 3, 1, 2).contiguous().cpu().numpy(), rois[pick]).transpose(0, 2, 3, 1)
    assert np.array_equal(got_b[torch.from_numpy(pick).to(dev)].float().cpu().numpy(), ref_b)
    # properties over all 20 000 rows, on the device
    rev = ops.roi_pool(fb, torch.flip(r, dims=[0]).contiguous(), layout="NHWC")
    assert torch.equal(torch.flip(rev, dims=[0]).view(torch.int16), got_b.view(torch.int16))
    ch_max = fb.view(-1, C).float().amax(dim=0)
    ch_min = fb.view(-1, C).float().amin(dim=0)
    v = got_b.view(-1, C).float()
    assert bool(((v <= ch_max) & ((v >= ch_min) | (v == 0))).all())
    del got, got_b, rev, v


def test_detection_step_equals_reference_golden(dev, golden):
    """The device detection step (azn_detect_rois / azn_detect_select / azn_detect_thresholds / azn_detect_filter /
    azn_nms_segments) == the reference's own test_net (detections.pkl) and apply_nms, driven by the same HashDetNet:
    counts, scores and order bit for bit, boxes to 1e-5 relative (float32 exp), NMS keeps identical."""
    from aznet_b200 import detector
    from oracle.gen_golden import DETECT_CASES, detect_case_proposals
    g = golden["detect"]

    class Head:                                              # dimensions only: head outputs are supplied
        pooled, C = 7, 8

        def __init__(self, ncls):
            self.num_classes = ncls
            self.w6 = torch.zeros((64, 7 * 7 * 8), dtype=torch.bfloat16, device=dev)
            self.w7 = torch.zeros((64, 64), dtype=torch.bfloat16, device=dev)

    for name, ncls, shapes, max_size, bs, counts, nms_t in DETECT_CASES:
        n_img, cap = len(shapes), 300
        props = detect_case_proposals(shapes, counts)
        net = synth.HashDetNet(seed=13, num_classes=ncls)
        dset = detector.DetectionSet(n_img, ncls, device=dev)
        for shape in sorted(set(shapes)):                    # one engine per image shape, slices of the set-wide buffers
            idx = [i for i, s in enumerate(shapes) if s == shape]
            eng = detector.DetectEngine(Head(ncls), len(idx), shape[0], shape[1], cap, max_size=max_size, batch_size=bs)
            boxes = np.zeros((len(idx), cap, 4))
            for k, i in enumerate(idx):
                boxes[k, :counts[i]] = props[i]
            eng.prepare(torch.from_numpy(boxes).to(dev), torch.tensor([counts[i] for i in idx], dtype=torch.int32, device=dev))
            m = int(eng.m_total.item())
            rois = eng.rois[:m].cpu().numpy().copy()
            rois[:, 0] = 0
            p, d = net.heads(rois)
            head = np.zeros((len(idx) * cap, eng.ld), np.float32)
            head[:m, :ncls], head[:m, ncls:5 * ncls] = p, d
            dets, tops, cnt = eng.select(head_out=torch.from_numpy(head).to(dev))
            for k, i in enumerate(idx):
                dset.dets[i], dset.top_scores[i], dset.det_count[i] = dets[k], tops[k], cnt[k]
        dset.finish(nms_t)
        for tag, ab in (("det", dset.to_host()), ("nms", dset.to_host(nms=True))):
            rows, rc = g["%s_%s_rows" % (name, tag)], g["%s_%s_count" % (name, tag)]
            pos = 0
            for j in range(ncls):
                for i in range(n_img):
                    n = int(rc[j, i])
                    got = ab[j][i]
                    assert len(got) == n, (name, tag, j, i, len(got), n)
                    if n:
                        ref = rows[pos:pos + n]
                        assert np.array_equal(got[:, 4].view(np.uint32), ref[:, 4].view(np.uint32)), (name, tag, j, i)
                        np.testing.assert_allclose(got[:, :4], ref[:, :4], rtol=1e-5, atol=1e-4)
                    pos += n
            assert pos == rows.shape[0]


def _iou(a, b):
    x1, y1 = np.maximum(a[:, None, 0], b[None, :, 0]), np.maximum(a[:, None, 1], b[None, :, 1])
    x2, y2 = np.minimum(a[:, None, 2], b[None, :, 2]), np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1 + 1, 0, None) * np.clip(y2 - y1 + 1, 0, None)
    aa = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
    ab = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    return inter / (aa[:, None] + ab[None, :] - inter)


@pytest.mark.parametrize("max_size", [800, 1000], ids=["voc_30x50", "default_38x63"])
def test_full_width_search_vs_oracle(dev, O, max_size):
    """The headline configuration at full width: 600x1000 images, 512-channel conv5_3 (30x50 with voc.yml's MAX_SIZE
    800, 38x63 with the default cfg), int6 25088 -> 4096, int7 1280, 56 head columns, the bench's seed-3 weights.
    Stated tolerance (bf16 operands and activations vs the fp32 oracle on the same bf16-rounded weights and maps):
    regions evaluated within 3 % (+-2), proposal recall at IoU >= 0.9 >= 0.95 both ways, top-20 scores within 3e-2."""
    from aznet_b200 import engine, ops
    import bench
    H, W, n_img = 600, 1000, 2
    w = synth.make_az_weights(seed=3, zoom_bias=bench.ZOOM_BIAS)
    s = engine.im_scale_for(H, W, (600,), max_size)
    fh, fw = synth.conv_shape(H, W, s)
    assert (fh, fw) == ((30, 50) if max_size == 800 else (38, 63))
    conv = synth.make_conv_maps(n_img, 512, fh, fw, seed=107)
    bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    wq = {k: (bf(v[0]), v[1]) for k, v in w.items()}
    head = engine.AZHeadWeights(w, dev)
    cfgk = dict(bench.CFG, max_size=max_size)
    eng = engine.SearchEngine(head, n_img, H, W, **cfgk)
    eng.propose(ops.nchw_to_nhwc_bf16(torch.from_numpy(conv).to(dev)))
    boxes, scores, n_eval, depth = eng.results()
    cfg = O.OracleCfg(TEST_MAX_SIZE=max_size, Tz=cfgk["tz"], NUM_PROPOSALS=cfgk["num_proposals"], BATCH_SIZE=cfgk["batch_size"])
    net = O.OracleNet(wq, "az", cfg=cfg, act_round=O.round_bf16)
    for i in range(n_img):
        Y, sc, info = O.im_propose({"full": net, "fc": net}, (H, W, 3), cfg, conv={"conv5_3": bf(conv[i:i + 1])}, return_scores=True)
        assert abs(int(n_eval[i]) - info["num_eval"]) <= max(2, 0.03 * info["num_eval"]), (int(n_eval[i]), info["num_eval"])
        m = _iou(Y, boxes[i])
        assert (m.max(1) >= 0.9).mean() >= 0.95 and (m.max(0) >= 0.9).mean() >= 0.95
        top = min(20, len(sc), len(scores[i]))
        np.testing.assert_allclose(np.sort(scores[i])[::-1][:top], np.sort(sc)[::-1][:top], atol=3e-2)


def test_fc_forward_int6_shape_of_the_deepest_level(dev):
    """fc_gemm<256,2> at the bench's deepest-level shape (M = 1494 live rows of 1536, N = 4096, K = 25088) vs a plain
    fp32 product of the same bf16 operands: |err| <= 3e-2 on O(1) outputs (bf16 output rounding 2^-8 relative)."""
    from aznet_b200 import _lib, ops
    M, Mcap, N, K = 1494, 1536, 4096, 25088
    gen = torch.Generator(device="cpu").manual_seed(5)
    x = torch.relu(torch.randn((Mcap, K), generator=gen)).to(torch.bfloat16).to(dev)
    w = (torch.randn((N, K), generator=gen) * (2.0 / K) ** 0.5).to(torch.bfloat16).to(dev)
    b = (torch.randn((N,), generator=gen) * 0.01).to(dev)
    out = torch.full((Mcap, N), -7.0, dtype=torch.bfloat16, device=dev)
    ops.fc_forward(x, w, b, _lib.ACT_RELU, m_live=torch.tensor([M], dtype=torch.int32, device=dev), out=out)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.relu(x[:M].float() @ w.float().t() + b)
    err = (out[:M].float() - ref).abs().max().item()
    assert err <= 3e-2, err
    assert torch.all(out[M:] == -7.0)                        # rows past the live count are not written


def test_nms_20000_at_0p7_equals_oracle(dev, O):
    """BASELINE config #4's largest NMS case (N = 20000, IoU 0.7): keep list identical to the oracle of nms.pyx."""
    from aznet_b200 import ops
    d = synth.make_dets(20000, 600, 1000, seed=3)
    keep, cnt = ops.nms(torch.from_numpy(d).to(dev), 0.7)
    ref = O.nms(d, 0.7)
    assert 12000 < len(ref) < 14000                           # ~64 % kept at IoU 0.7 with this generator (12760 with its own)
    assert keep[:int(cnt.item())].cpu().tolist() == ref
