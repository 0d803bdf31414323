"""GPU suite: the skip-layer detector head (models/COCO/VGG16_skip/frcnn/test_fc.prototxt:28-232,
experiments/cfgs/voc_skip.yml) -- GRN + concat + x1000 (azn_grn_concat_forward), conv_pool5 as a tensor-core GEMM
over pooled positions, and the drop-in nets around them -- against the oracle's restatement of grn_layer.cpp."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from aznet_b200 import synth  # noqa: E402
from helpers import RowMatcher, _same_blob_for_both_routes  # noqa: E402

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


def _bf(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).float().numpy()


@pytest.mark.parametrize("channels", [(32, 64, 64), (256, 512, 512), (8,), (1024, 40)])
def test_grn_concat_matches_oracle(dev, O, channels):
    """out[r, off_l + c] = bf16(1000 * x / sqrt(sum_c x^2)) per source; fp32 summation order is the only freedom,
    so the result is within one bf16 ulp of the oracle's grn(); live-row count and padded columns are respected."""
    from aznet_b200 import ops
    rng = np.random.default_rng(3)
    rows, live_units, per = 49 * 40, 33, 49
    srcs = [_bf(np.maximum(rng.standard_normal((rows, c), dtype=np.float32), 0) * np.float32(3.0)) for c in channels]
    srcs[0][7] = 0.0                                           # an all-zero position: 0/0 -> NaN like the reference (no epsilon)
    ctot = sum(channels)
    ld = (ctot + 63) // 64 * 64
    out = torch.full((rows, ld), 7.0, dtype=torch.bfloat16, device=dev)
    n_units = torch.tensor([live_units], dtype=torch.int32, device=dev)
    ops.grn_concat([torch.from_numpy(s).to(dev).to(torch.bfloat16) for s in srcs], 1000.0, n_units=n_units, rows_per_unit=per, out=out)
    got = out.float().cpu().numpy()
    live = live_units * per
    assert np.all(got[live:] == 7.0) and np.all(got[:, ctot:] == 7.0), "rows past the live count / pad columns were written"
    off = 0
    for s, c in zip(srcs, channels):
        ref = O.grn(s[:live].reshape(live, c, 1, 1)).reshape(live, c) * np.float32(1000.0)
        g = got[:live, off:off + c]
        assert np.all(np.isnan(g[7]) == np.isnan(ref[7])) and (off > 0 or np.all(np.isnan(g[7])))
        ok = ~np.isnan(ref)
        np.testing.assert_allclose(g[ok], ref[ok], rtol=2.0 ** -7, atol=1e-6)
        off += c
    # without a device count every row is processed
    out2 = ops.grn_concat([torch.from_numpy(s).to(dev).to(torch.bfloat16) for s in srcs], 1000.0)
    assert out2.shape == (rows, ctot)
    assert torch.equal(out2[:live].view(torch.int16), out[:live, :ctot].contiguous().view(torch.int16))


def _interior_rois(n, H, W, seed):
    """ROIs whose bins are all non-empty at every pooling scale (an empty bin is an all-zero position -> NaN by GRN)."""
    b = synth.make_boxes(n, H - 40, W - 40, seed=seed, lo=40, hi=180)
    return np.hstack([np.zeros((n, 1)), b + 8.0]).astype(np.float32)


def _skip_setup(dev, num_classes=6):
    from aznet_b200 import backbone, net
    bw = backbone.make_vgg16_weights(seed=5, width_div=8)              # conv3_3: 32, conv4_3 / conv5_3: 64 channels
    bb = backbone.VGG16Backbone(bw, dev)
    w = synth.make_frcnn_skip_weights(seed=4, num_classes=num_classes, channels=(32, 64, 64), c_out=64, h6=256, h7=128)
    w["cls_score"] = (w["cls_score"][0] * np.float32(0.2), w["cls_score"][1])
    w["bbox_pred"] = (w["bbox_pred"][0] * np.float32(0.5), w["bbox_pred"][1])
    nets = {"full": net.Net(w, "frcnn_skip", backbone=bb, name="skip_small"), "fc": net.Net(w, "frcnn_skip", name="skip_small")}
    return nets, w, bb


def test_skip_head_forward_vs_oracle(dev, O):
    """Net(kind='frcnn_skip').forward(conv3_3, conv4_3, conv5_3, rois) against the oracle net on the same maps and
    bf16-rounded weights, with the product's bf16 storage of the concat operand emulated."""
    nets, w, _ = _skip_setup(dev)
    rng = np.random.default_rng(11)
    H, W = 240, 320
    maps = {"conv3_3": _bf(np.maximum(rng.standard_normal((1, 32, H // 4, W // 4), dtype=np.float32), 0)),
            "conv4_3": _bf(np.maximum(rng.standard_normal((1, 64, H // 8, W // 8), dtype=np.float32), 0)),
            "conv5_3": _bf(np.maximum(rng.standard_normal((1, 64, H // 16, W // 16), dtype=np.float32), 0))}
    rois = _interior_rois(150, H, W, seed=9)
    fc = nets["fc"]
    assert fc.inputs == ["conv3_3", "conv4_3", "conv5_3", "rois"]
    with pytest.raises(Exception, match="do not match net inputs"):
        fc.forward(rois=rois, conv5_3=maps["conv5_3"])
    for n, m in maps.items():
        fc.blobs[n].reshape(*m.shape)
    fc.blobs["rois"].reshape(*rois.shape)
    out = fc.forward(rois=rois, **maps)
    wq = {k: (_bf(v[0]), v[1]) for k, v in w.items() if k != "skip_channels"}
    onet = O.OracleNet(wq, "frcnn_skip", act_round=O.round_bf16)
    for n, m in maps.items():
        onet.blobs[n].reshape(*m.shape)
    onet.blobs["rois"].reshape(*rois.shape)
    ref = onet.forward(rois=rois, **maps)
    assert out["cls_prob"].shape == (150, 6) and out["bbox_pred"].shape == (150, 24)
    assert np.isfinite(ref["cls_prob"]).all() and np.isfinite(out["cls_prob"]).all()
    np.testing.assert_allclose(out["cls_prob"], ref["cls_prob"], atol=1e-2)
    np.testing.assert_allclose(out["bbox_pred"], ref["bbox_pred"], atol=2e-2)
    # pool5 itself (GRN + concat + conv_pool5): O(1) activations, bf16 output rounding
    p5 = fc.skip_pool5({n: fc._resident_map(n, m) for n, m in maps.items()}, torch.from_numpy(rois).to(dev))
    p5 = p5.float().cpu().numpy().reshape(150, 7, 7, 64).transpose(0, 3, 1, 2)
    p5_ref = O.skip_pool5(wq, maps, rois, act_round=O.round_bf16)
    np.testing.assert_allclose(p5, p5_ref, rtol=2e-2, atol=2e-2)


def test_skip_empty_bin_is_nan_like_the_reference(dev, O):
    """A ROI hugging the right border has empty bins -> all-zero positions -> GRN divides 0 by 0: the reference's
    scores for that ROI are NaN (and never pass `score > thresh`); so are ours (NaN-propagating ReLU in the GEMM)."""
    nets, w, _ = _skip_setup(dev)
    rng = np.random.default_rng(12)
    H, W = 240, 320
    maps = {"conv3_3": _bf(np.maximum(rng.standard_normal((1, 32, H // 4, W // 4), dtype=np.float32), 0)),
            "conv4_3": _bf(np.maximum(rng.standard_normal((1, 64, H // 8, W // 8), dtype=np.float32), 0)),
            "conv5_3": _bf(np.maximum(rng.standard_normal((1, 64, H // 16, W // 16), dtype=np.float32), 0))}
    rois = np.array([[0, 310, 100, 319, 140], [0, 50, 50, 150, 150]], dtype=np.float32)
    fc = nets["fc"]
    for n, m in maps.items():
        fc.blobs[n].reshape(*m.shape)
    fc.blobs["rois"].reshape(*rois.shape)
    out = fc.forward(rois=rois, **maps)
    wq = {k: (_bf(v[0]), v[1]) for k, v in w.items() if k != "skip_channels"}
    onet = O.OracleNet(wq, "frcnn_skip")
    for n, m in maps.items():
        onet.blobs[n].reshape(*m.shape)
    onet.blobs["rois"].reshape(*rois.shape)
    ref = onet.forward(rois=rois, **maps)
    assert np.isnan(ref["cls_prob"][0]).all() and np.isnan(out["cls_prob"][0]).all()
    assert np.isfinite(ref["cls_prob"][1]).all() and np.isfinite(out["cls_prob"][1]).all()


def test_skip_shared_detection_dropin(dev, O, capsys, monkeypatch):
    """voc_skip.yml wiring: SEAR.FRCNN_CONV = [conv3_3, conv4_3, conv5_3], DEDUP_BOXES = 0.5.  im_detect_shared hands
    the three maps of the AZ-Net pass to the skip detector; the unshared im_detect recomputes them from the image;
    both agree, and both agree with the host-route over foreign (wrapped) nets."""
    from aznet_b200 import net
    from aznet_b200.detect import config as C
    from aznet_b200.detect import test as T
    _same_blob_for_both_routes(monkeypatch, dev)
    nets, w, bb = _skip_setup(dev)
    azw = synth.make_az_weights(seed=3, C=64, h6=256, h71=96, h72=32, zoom_bias=0.0)
    az = {"full": net.Net(azw, "az", backbone=bb, name="az_small"), "fc": net.Net(azw, "az", name="az_small")}
    cfg = C.cfg
    saved = (list(cfg.SEAR.FRCNN_CONV), cfg.DEDUP_BOXES)
    cfg.SEAR.FRCNN_CONV, cfg.DEDUP_BOXES = ["conv3_3", "conv4_3", "conv5_3"], 0.5
    C.cfg_set_mode("Test", 0.5)
    try:
        im = synth.make_images(1, 240, 320, seed=3)[0]
        s_shared, p_shared = T.im_detect_shared(az, nets, im, 6)
        boxes = T.im_propose(az, im)
        s_full, p_full = T.im_detect(nets, im, boxes, 6)
        assert s_shared.shape == (boxes.shape[0], 6) and p_shared.shape == (boxes.shape[0], 24)
        ok = np.isfinite(s_full).all(axis=1)
        assert ok.mean() > 0.5
        assert np.array_equal(np.isfinite(s_shared).all(axis=1), ok)
        np.testing.assert_allclose(s_shared[ok], s_full[ok], atol=1e-5)
        np.testing.assert_allclose(p_shared[ok], p_full[ok], rtol=1e-5, atol=1e-4)
        # host route under the skip config: the AZ 'full' net is asked for (and caches) all of SEAR.FRCNN_CONV like the
        # reference (test.py:222-226), so the skip detector finds conv3_3 / conv4_3 in the shared dict
        class Foreign(dict):
            pass
        wrap = lambda n: type("W", (), {"forward": n.forward, "blobs": n.blobs, "name": n.name})()
        s_host, p_host = T.im_detect_shared(Foreign(full=wrap(az["full"]), fc=wrap(az["fc"])),
                                            Foreign(full=wrap(nets["full"]), fc=wrap(nets["fc"])), im, 6)
        assert s_host.shape[1] == 6 and p_host.shape == (s_host.shape[0], 24) and abs(s_host.shape[0] - s_shared.shape[0]) <= 3
        assert np.isfinite(s_host).all(axis=1).mean() > 0.5
        # conv dict contract of im_propose(return_conv=True) under the skip config
        _, conv = T.im_propose(az, im, return_conv=True)
        assert sorted(conv) == ["conv3_3", "conv4_3", "conv5_3"]
        # 240x320 image, TEST.SCALES 600 -> 600x800 network input; ceil-mode pools: /4, /8, /16
        assert conv["conv3_3"].shape == (1, 32, 150, 200) and conv["conv4_3"].shape == (1, 64, 75, 100) and conv["conv5_3"].shape == (1, 64, 38, 50)
        assert all(v.dtype == np.float32 for v in conv.values())
    finally:
        cfg.SEAR.FRCNN_CONV, cfg.DEDUP_BOXES = saved
        C.cfg_set_mode("Test", 0.5)
    capsys.readouterr()


def test_skip_drivers_device_route_equals_host_route(dev, tmp_path, capsys, monkeypatch):
    """test_net / test_net_shared with the skip-layer detector: DetectEngine's skip head (three staged/direct ROI pools,
    azn_grn_concat_forward, conv_pool5 GEMM with a device-side live row count) against the reference's host loop over
    the same Net objects."""
    import os
    import pickle
    import cv2
    from aznet_b200 import net
    from aznet_b200.detect import config as C
    from aznet_b200.detect import test as T
    _same_blob_for_both_routes(monkeypatch, dev)
    nets, w, bb = _skip_setup(dev)
    azw = synth.make_az_weights(seed=3, C=64, h6=256, h71=96, h72=32, zoom_bias=0.0)
    az = {"full": net.Net(azw, "az", backbone=bb, name="az_small"), "fc": net.Net(azw, "az", name="az_small")}
    cfg = C.cfg
    saved = (list(cfg.SEAR.FRCNN_CONV), cfg.DEDUP_BOXES, cfg.ROOT_DIR)
    cfg.SEAR.FRCNN_CONV, cfg.DEDUP_BOXES = ["conv3_3", "conv4_3", "conv5_3"], 0.5
    C.cfg_set_path("pytest_skip")
    C.cfg_set_mode("Test", 0.5)
    try:
        paths = []
        for i, im in enumerate(synth.make_images(3, 200, 300, seed=70)):
            p = str(tmp_path / ("im%d.png" % i))
            cv2.imwrite(p, im)
            paths.append(p)
        cfg.ROOT_DIR = str(tmp_path)
        imdb = synth.SyntheticImdb(paths, num_classes=6)
        T.test_proposals(az, imdb)
        prop_file = os.path.join(C.get_output_dir(imdb, az["full"]), "proposals.pkl")
        det_file = os.path.join(C.get_output_dir(imdb, nets["full"]), "detections.pkl")

        class Foreign(dict):
            pass
        wrap = lambda n: type("W", (), {"forward": n.forward, "blobs": n.blobs, "name": n.name})()

        def run(dn, shared):
            if shared:
                T.test_net_shared(az, dn, imdb)
            else:
                T.test_net(dn, prop_file, imdb)
            capsys.readouterr()
            return pickle.load(open(det_file, "rb")), imdb.evaluated[0]

        total, match = 0, RowMatcher(atol=0.5)
        for shared in (False, True):
            pre_d, nms_d = run(nets, shared)
            pre_h, nms_h = run(Foreign(full=wrap(nets["full"]), fc=wrap(nets["fc"])), shared)
            for a_set, b_set in ((pre_d, pre_h), (nms_d, nms_h)):
                for j in range(1, 6):
                    for i in range(3):
                        a, b = a_set[j][i], b_set[j][i]
                        total += len(a)
                        match.add(a, b)      # 3 images per GEMM on the batched route, one on the host loop
        assert total > 50
        match.check(0.95)
    finally:
        cfg.SEAR.FRCNN_CONV, cfg.DEDUP_BOXES, cfg.ROOT_DIR = saved
        C.cfg_set_mode("Test", 0.5)


def test_roi_pool_grn_fused_equals_pool_then_grn(dev):
    """azn_roi_pool_grn_fwd = azn_roi_pool_fwd followed by azn_grn_concat_forward, bit for bit (same summation order),
    for every source geometry of the skip head, with a device-side ROI count, empty bins (NaN) included."""
    from aznet_b200 import ops
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    n_img, H, W = 3, 240, 320
    rois = torch.from_numpy(synth.make_rois(500, H, W, seed=4, n_img=n_img)).to(dev)
    n_live = torch.tensor([437], dtype=torch.int32, device=dev)
    specs = [(32, 4, 0.25), (64, 8, 0.125), (512, 16, 0.0625), (256, 4, 0.25), (1024, 16, 0.0625)]
    ctot = sum(c for c, _, _ in specs)
    fused = torch.full((500 * 49, ctot + 24), 3.0, dtype=torch.bfloat16, device=dev)
    pooled, off = [], 0
    for c, stride, sc in specs:
        m = torch.randn((n_img, H // stride, W // stride, c), generator=g, device=dev).clamp_(min=0).to(torch.bfloat16).contiguous()
        ops.roi_pool_grn(m, rois, fused, off, 7, sc, 1000.0, n_rois=n_live)
        pooled.append(ops.roi_pool(m, rois, 7, sc, layout="NHWC", n_rois=n_live).view(500 * 49, c))
        off += c
    ref = torch.full((500 * 49, ctot + 24), 3.0, dtype=torch.bfloat16, device=dev)
    ops.grn_concat(pooled, 1000.0, n_units=n_live, rows_per_unit=49, out=ref)
    assert torch.equal(fused.view(torch.int16), ref.view(torch.int16))
    live = fused[:437 * 49, :ctot].float()
    assert torch.isnan(live).any() and torch.isfinite(live).float().mean() > 0.9      # border ROIs have empty bins
    assert bool((fused[437 * 49:] == 3.0).all())
