"""CPU suite: pins the oracle restatement (oracle/) against the golden vectors produced by the
reference's own code (oracle/gen_golden.py) and, when oracle/_ref holds the compiled reference
Cython modules, against those modules directly on fresh random inputs."""
import numpy as np
import pytest

from aznet_b200 import synth
from oracle import az_oracle as O
from oracle import build_ref

SURVEY_ROOT_CHILDREN = np.array([                       # SURVEY.md section 8c golden vector
    [0., 0., 1000 / 3., 300.], [1000 / 3., 0., 2000 / 3., 300.], [2000 / 3., 0., 1000., 300.],
    [500 / 3., 150., 500., 450.], [500., 150., 2500 / 3., 450.],
    [0., 300., 1000 / 3., 600.], [1000 / 3., 300., 2000 / 3., 600.], [2000 / 3., 300., 1000., 600.]])


def test_divide_region_survey_vector():
    out = O.divide_region(np.array([[0, 0, 999, 599.]]), 10.0)
    np.testing.assert_allclose(out, SURVEY_ROOT_CHILDREN, rtol=0, atol=1e-9)


def test_divide_region_golden(golden):
    g = golden["div"]
    for k in g.files:
        if k.startswith("in_"):
            out = O.divide_region(g[k], 10.0)
            assert np.array_equal(out, g["out_" + k[3:]]), k
    lv = g["in_root_600x1000"]
    for lvl in range(2, 6):
        lv = O.divide_region(lv, 10.0)
        assert np.array_equal(lv, g["cascade_%d" % lvl])
    assert list(g["cascade_sizes"]) == [1, 8, 32, 134, 564]
    assert np.array_equal(O.sift_dup(g["sift_in"], 10.0), g["sift_out"])


def test_divide_region_empty():
    assert O.divide_region(np.zeros((0, 4)), 10.0).shape == (0, 4)


def test_nms_golden(golden):
    g = golden["nms"]
    for n in (1, 2, 17, 300, 2000):
        dets = synth.make_dets(n, seed=3)
        for th in (0.3, 0.5, 0.7):
            keep = O.nms(dets, th)
            assert keep == list(g["keep_n%d_t%d" % (n, int(th * 10))]), (n, th)
    assert O.nms(g["edge_dets"], 0.5) == list(g["edge_keep_t5"])
    assert O.nms(g["edge_dets"], 0.51) == list(g["edge_keep_t51"])


def test_nms_argument_checks():
    d = synth.make_dets(5)
    with pytest.raises(ValueError):
        O.nms(d.astype(np.float64), 0.5)
    with pytest.raises(TypeError):
        O.nms(d, np.float32(0.5))
    assert O.nms(np.zeros((0, 5), np.float32), 0.5) == []


@pytest.mark.skipif(not build_ref.have_ref_cython(), reason="oracle/_ref not built")
def test_against_compiled_reference_cython():
    div, nms, _ = build_ref.import_ref_cython()
    rng = np.random.default_rng(77)
    for t in range(5):
        b = synth.make_boxes(150, 480, 640, seed=100 + t, lo=11, hi=300)
        assert np.array_equal(O.divide_region(b, 10.0), div.divide_region(b, 10.0))
        d = synth.make_dets(700, 480, 640, seed=200 + t)
        th = float(rng.uniform(0.2, 0.8))
        assert O.nms(d, th) == nms.nms(d, th)


def _run_search_case(name, g):
    H, W, max_size, bs, tz, rate, nprop, fixed = g[name + "_cfg"]
    cfg = O.OracleCfg(TEST_MAX_SIZE=int(max_size), BATCH_SIZE=int(bs), Tz=float(tz),
                      FIXED_PROPOSAL_NUM=bool(fixed), NUM_PROPOSALS=300 if nprop < 0 else int(nprop),
                      APPEND_BOXES=name.startswith("append"))
    net = synth.HashNet(seed=11, zoom_rate=float(rate))
    conv = {"conv5_3": np.zeros((1, 1, 2, 2), np.float32)}
    Y, s, info = O.im_propose({"full": net, "fc": net}, (int(H), int(W), 3), cfg, conv=conv, return_scores=True)
    return Y, info


@pytest.mark.parametrize("name", ["d0_600x1000", "voc_600x1000", "fullzoom_600x1000", "small_375x500",
                                  "chunked_480x640", "tc_thresh_333x500", "nozoom_600x1000", "append_375x500"])
def test_im_propose_golden(golden, name):
    g = golden["search"]
    Y, info = _run_search_case(name, g)
    ref = g[name + "_Y"]
    log = str(g[name + "_log"])
    assert "{0} proposals, evaluate {1} regions, reaches depth {2}.".format(
        Y.shape[0], info["num_eval"], info["depth"]) == log
    # HashNet scores are 24-bit uniform draws: ties are possible but rare; the reference used an
    # unstable argsort, so compare as sets of rows when the ordered compare fails.
    if not np.array_equal(Y, ref):
        assert sorted(map(tuple, Y)) == sorted(map(tuple, ref))


TUNE_CASES = ["train_375x500", "train_voc_600x1000", "tz05_480x640", "stop_early_333x500"]


def _run_tune_case(name, g):
    H, W, max_size, bs, tz, rate, nprop = g[name + "_cfg"]
    cfg = O.OracleCfg(TEST_MAX_SIZE=int(max_size), BATCH_SIZE=int(bs), Tz=float(tz), NUM_PROPOSALS=int(nprop))
    net = synth.HashNet(seed=11, zoom_rate=float(rate))
    conv = {"conv5_3": np.zeros((1, 1, 2, 2), np.float32)}
    return O.im_propose_tune({"full": net, "fc": net}, (int(H), int(W), 3), cfg, conv=conv)


@pytest.mark.parametrize("name", TUNE_CASES)
def test_im_propose_tune_golden(golden, name):
    """The restated diagnostic search equals the reference's own lib/detect/tune.py run (history bit for bit)."""
    g = golden["tune"]
    Y5, Bhis, info = _run_tune_case(name, g)
    assert "{0} proposals, evaluate {1} regions, reaches depth {2}.".format(
        Y5.shape[0], info["num_eval"], info["depth"]) == str(g[name + "_log"])
    assert np.array_equal(Bhis.view(np.uint64), g[name + "_Bhis"].view(np.uint64))
    ref = g[name + "_Y5"]
    if not np.array_equal(Y5, ref):
        assert sorted(map(tuple, Y5)) == sorted(map(tuple, ref))


def test_tune_thresh_golden(golden):
    """tune_thresh's heap over six images (three sizes) = the reference's thresh.pkl."""
    g = golden["tune"]
    cfg = O.OracleCfg(Tz=0.0, NUM_PROPOSALS=2000)
    net = synth.HashNet(seed=11, zoom_rate=0.5)
    conv = {"conv5_3": np.zeros((1, 1, 2, 2), np.float32)}
    hist = [O.im_propose_tune({"full": net, "fc": net}, (int(h), int(w), 3), cfg, conv=conv)[1] for h, w in g["thresh_shapes"]]
    assert O.tune_thresh(hist, 20) == float(g["thresh_per20"])
    assert O.tune_thresh(hist, 1000000) == -np.inf
    # order-independence, which is what lets the product compute it in one batched device pass
    assert O.tune_thresh(hist[::-1], 20) == float(g["thresh_per20"])
    allz = np.sort(np.concatenate([h[:, 4] for h in hist]))[::-1]
    assert allz[20 * len(hist) - 1] == float(g["thresh_per20"])


def test_decode_clip_unwrap_golden(golden):
    g = golden["search"]
    pred = O.bbox_pred(g["bbox_boxes"], g["bbox_deltas"])
    assert np.array_equal(pred, g["bbox_pred"])
    clip = O.clip_boxes(pred.copy(), (600, 1000, 3))
    assert np.array_equal(clip, g["bbox_clip"])
    a, c = O.unwrap_adj_pred(clip, g["unwrap_scores_in"], 10)
    assert np.array_equal(a, g["unwrap_boxes"]) and np.array_equal(c, g["unwrap_scores"])
    assert O.bbox_pred(np.zeros((0, 4)), np.zeros((0, 44), np.float32)).shape == (0, 44)


def test_roi_pool_matches_torchvision_cpu():
    """Secondary cross-check (the reference pins no forward values for ROIPooling)."""
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    feat = synth.make_conv_maps(2, 16, 38, 63, seed=7)
    rois = synth.make_rois(300, 600, 1000, seed=5, n_img=2)
    rois[:10, 1:] = np.round(rois[:10, 1:] / 8) * 8          # coordinates landing on .5 after *1/16
    rois[10:14, 1:] += 900                                   # out-of-image ROIs -> empty bins
    rois[14] = [0, 50, 50, 40, 40]                           # malformed (end < start) -> 1x1
    out = O.roi_pool_fwd(feat, rois)
    ref = tv.ops.roi_pool(torch.from_numpy(feat), torch.from_numpy(rois), (7, 7), 0.0625).numpy()
    assert np.array_equal(out, ref)


def _layer_feat():
    feat = synth.make_conv_maps(2, 8, 38, 63, seed=7)          # as oracle/gen_golden.py::gen_layers
    feat[1] -= 0.5
    feat[1, 0, 3, 4] = np.nan
    feat[1, 1, 10:20, 10:30] = -0.0
    return feat


def test_caffe_layers_golden(golden):
    """The restated layers equal, BIT FOR BIT, what Forward_cpu of the reference's own unmodified layer sources
    produced (oracle/_ref/libcaffe_layers_ref.so via oracle/gen_golden.py --only layers): ROI max-pool incl. argmax on
    the edge-ROI set and the 739 natural regions (negative values, NaN, -0 in the map), GRN incl. the 0/0 position,
    sigmoid, ReLU, softmax, MAX pooling in ceil mode; InnerProduct within fp32 summation-order noise."""
    g = golden["caffe_layers"]
    feat = _layer_feat()
    for tag in ("edge", "natural"):
        out, am = O.roi_pool_fwd(feat, g["roi_%s_rois" % tag], want_argmax=True)
        assert np.array_equal(out.view(np.uint32), g["roi_%s_out" % tag].view(np.uint32)), tag
        assert np.array_equal(am, g["roi_%s_argmax" % tag].astype(np.int32)), tag
    assert g["roi_natural_rois"].shape[0] == 739
    with np.errstate(all="ignore"):
        assert np.array_equal(O.grn(g["grn_in"]).view(np.uint32), g["grn_out"].view(np.uint32))
    assert np.isnan(g["grn_out"][0, :, 0, 0]).all()
    assert np.array_equal(O.sigmoid(g["sigmoid_in"]), g["sigmoid_out"])
    assert np.array_equal(O.relu(g["sigmoid_in"]), g["relu_out"])
    assert np.array_equal(O.softmax(g["softmax_in"]), g["softmax_out"])
    x = g["ip_x"].reshape(g["ip_x"].shape[0], -1)               # axis-1 flattening = c*49 + ph*7 + pw
    np.testing.assert_allclose(O.inner_product(x, g["ip_w"], g["ip_b"]), g["ip_out"], rtol=0, atol=2e-5)
    assert np.array_equal(O.max_pool_ceil(g["maxpool_in"]), g["maxpool_out"])


def test_against_compiled_reference_layers():
    """Fresh random inputs through the compiled reference layer sources (authoring container and GPU box: the .so
    travels prebuilt)."""
    from oracle import ref_caffe as RC
    if not RC.available():
        pytest.skip("oracle/_ref/libcaffe_layers_ref.so not built")
    rng = np.random.default_rng(123)
    for t in range(3):
        C, H, W = int(rng.integers(1, 9)), int(rng.integers(5, 40)), int(rng.integers(5, 64))
        feat = rng.standard_normal((2, C, H, W)).astype(np.float32)
        rois = synth.make_rois(400, H * 16, W * 16, seed=50 + t, n_img=2)
        rois[:40, 1:] = np.round(rois[:40, 1:] / 8) * 8
        rois[40:60, 1:] -= 100
        for pooled, sc in ((7, 0.0625), (6, 0.125), (3, 0.25)):
            a, am = O.roi_pool_fwd(feat, rois, pooled, sc, want_argmax=True)
            b, bm = RC.roi_pool_fwd(feat, rois, pooled, sc, want_argmax=True)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(am, bm)
        x = rng.standard_normal((5, 16, 7, 7)).astype(np.float32)
        assert np.array_equal(O.grn(x), RC.grn(x))
        z = (rng.standard_normal((33, 21)) * 6).astype(np.float32)
        assert np.array_equal(O.softmax(z), RC.softmax(z)) and np.array_equal(O.sigmoid(z), RC.sigmoid(z))
        pm = rng.standard_normal((1, 2, int(rng.integers(3, 50)), int(rng.integers(3, 50)))).astype(np.float32)
        assert np.array_equal(O.max_pool_ceil(pm), RC.max_pool(pm))
    with pytest.raises(RuntimeError):
        RC.roi_pool_fwd(np.zeros((1, 1, 4, 4), np.float32), np.array([[2, 0, 0, 10, 10]], np.float32))


def test_reference_im_propose_with_a_real_net_equals_the_port():
    """The reference's OWN lib/detect/test.py::im_propose (oracle/_ref/pyref) + its own compiled div.pyx, driving its own
    layer sources (oracle/_ref/libcaffe_layers_ref.so: ROIPooling / ReLU / Sigmoid) with float32 sgemm heads -- a REAL
    network, not the integer-hash one -- gives bit for bit the proposals of the oracle port (which is what the CUDA
    path is compared with at full width).  This is also the stack bench.py's reference arm times."""
    from oracle import ref_caffe as RC
    ref = build_ref.load_pyref()
    if ref is None or not RC.available():
        pytest.skip("oracle/_ref not built")
    rtest, rconfig = ref[0], ref[1]
    w = synth.make_az_weights(seed=3, C=64, h6=256, h71=96, h72=32, zoom_bias=0.0)
    conv = synth.make_conv_maps(1, 64, 30, 50, seed=7)
    cfg = O.OracleCfg(TEST_MAX_SIZE=800, BATCH_SIZE=1000)
    port = O.OracleNet(w, "az", cfg=cfg)
    Y_port, _, info = O.im_propose({"full": port, "fc": port}, (600, 1000, 3), cfg, conv={"conv5_3": conv}, return_scores=True)
    rc = rconfig.cfg
    keep = (rc.TEST.MAX_SIZE, rc.SEAR.BATCH_SIZE)
    rc.TEST.MAX_SIZE, rc.SEAR.BATCH_SIZE = 800, 1000
    rconfig.cfg_set_mode("Test", 0.5)
    try:
        full = O.OracleNet(w, "az", cfg=cfg, layers="ref", backbone=lambda data: conv)      # the backbone is out of scope: cached map
        fc = O.OracleNet(w, "az", cfg=cfg, layers="ref")
        import contextlib
        import io
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            Y_ref = rtest.im_propose({"full": full, "fc": fc}, np.zeros((600, 1000, 3), np.uint8))
    finally:
        rc.TEST.MAX_SIZE, rc.SEAR.BATCH_SIZE = keep
    assert "evaluate {0} regions".format(info["num_eval"]) in buf.getvalue() and info["num_eval"] > 100
    assert np.array_equal(Y_ref.view(np.uint64), Y_port.view(np.uint64))


def _detect_case(g, name, ncls, shapes, max_size, bs, counts):
    from oracle.gen_golden import detect_case_proposals
    cfg = O.OracleCfg(TEST_MAX_SIZE=max_size, BATCH_SIZE=bs)
    net = synth.HashDetNet(seed=13, num_classes=ncls)
    props = detect_case_proposals(shapes, counts)
    per_image = []
    for (h, w), p in zip(shapes, props):
        if p.shape[0] == 0:
            per_image.append(None)
            continue
        s_ref, p_ref, _ = O.frcnn_forward({"full": net, "fc": net}, (h, w, 3), p, ncls,
                                          {"conv5_3": np.zeros((1, 1, 2, 2), np.float32)}, cfg)
        per_image.append((s_ref, p_ref))
    return per_image


def _flatten(all_boxes, ncls, n_img):
    rows, cnt = [], np.zeros((ncls, n_img), np.int32)
    for j in range(ncls):
        for i in range(n_img):
            d = all_boxes[j][i]
            if isinstance(d, list):
                continue
            cnt[j, i] = d.shape[0]
            rows.append(d)
    return (np.vstack(rows) if rows else np.zeros((0, 5), np.float32)), cnt


def test_detection_path_golden(golden):
    """frcnn_forward / test_net_select / apply_nms of the oracle equal the reference's OWN im_detect, test_net
    (detections.pkl) and apply_nms (what evaluate_detections received) bit for bit (oracle/gen_golden.py --only
    detect; HashDetNet scores are tie-free 24-bit draws per (roi, class))."""
    from oracle.gen_golden import DETECT_CASES
    g = golden["detect"]
    for name, ncls, shapes, max_size, bs, counts, nms_t in DETECT_CASES:
        per_image = _detect_case(g, name, ncls, shapes, max_size, bs, counts)
        i0 = int(g[name + "_imdet_index"])
        assert np.array_equal(per_image[i0][0], g[name + "_imdet_scores"].astype(np.float64)), name
        assert np.array_equal(per_image[i0][1].view(np.uint64), g[name + "_imdet_boxes"].view(np.uint64)), name
        all_boxes, thresh = O.test_net_select(per_image, ncls)
        rows, cnt = _flatten(all_boxes, ncls, len(shapes))
        assert np.array_equal(cnt, g[name + "_det_count"]), name
        assert np.array_equal(rows.view(np.uint32), g[name + "_det_rows"].view(np.uint32)), name
        rows, cnt = _flatten(O.apply_nms(all_boxes, nms_t), ncls, len(shapes))
        assert np.array_equal(cnt, g[name + "_nms_count"]), name
        assert np.array_equal(rows.view(np.uint32), g[name + "_nms_rows"].view(np.uint32)), name
        assert np.isfinite(thresh[1:]).any(), "the case must exercise the max_per_set threshold"


def test_roi_pool_argmax_and_errors():
    feat = synth.make_conv_maps(1, 4, 10, 12, seed=1)
    rois = np.array([[0, 0, 0, 191, 159], [0, 16, 16, 47, 47]], np.float32)
    out, am = O.roi_pool_fwd(feat, rois, want_argmax=True)
    flat = feat.reshape(1, 4, -1)
    sel = np.take_along_axis(np.broadcast_to(flat, (2, 4, 120)), am.reshape(2, 4, 49).astype(np.int64), axis=2)
    assert np.array_equal(sel.reshape(out.shape), out)
    with pytest.raises(RuntimeError):
        O.roi_pool_fwd(feat, np.array([[3, 0, 0, 10, 10]], np.float32))


def test_sigmoid_softmax_formulas():
    x = np.linspace(-20, 20, 4001).astype(np.float32)
    y = O.sigmoid(x)
    ref = (1.0 / (1.0 + np.exp(-x).astype(np.float64))).astype(np.float32)
    np.testing.assert_allclose(y, ref, rtol=4 * np.finfo(np.float32).eps)   # EXPECT_FLOAT_EQ = 4 ulp
    z = np.random.default_rng(0).standard_normal((50, 21)).astype(np.float32) * 5
    p = O.softmax(z)
    np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-5)
    e = np.exp(z - z.max(1, keepdims=True))
    np.testing.assert_allclose(p, e / e.sum(1, keepdims=True), atol=1e-4)


def test_test_net_select_small():
    rng = np.random.default_rng(5)
    C = 5
    per = []
    for i in range(6):
        s = rng.uniform(0, 1, (40, C)).astype(np.float32)
        b = rng.uniform(0, 100, (40, 4 * C))
        per.append((s, b) if i != 2 else None)
    all_boxes, thresh = O.test_net_select(per, C, max_per_image=10)
    max_per_set = 800 // (C - 1) * 6
    for j in range(1, C):
        tot = sum(len(all_boxes[j][i]) for i in range(6) if not isinstance(all_boxes[j][i], list))
        assert tot <= max_per_set
        assert all_boxes[j][2] == []


BLOB_TOL = 6.2e-5      # 4 float32 ulps at |v| < 256: cv2's SIMD path rounds the two multiply-adds differently


def test_image_blob_golden(golden_blob):
    """The numpy restatement of _get_image_blob vs the blobs the reference's own function produced
    (oracle/gen_golden.py --only blob)."""
    names = sorted(k[:-3] for k in golden_blob.files if k.endswith("_im"))
    assert len(names) >= 4
    for name in names:
        im, ref, c = golden_blob[name + "_im"], golden_blob[name + "_blob"], golden_blob[name + "_cfg"]
        cfg = O.OracleCfg(TEST_SCALES=(int(c[0]),), TEST_MAX_SIZE=int(c[1]))
        blob, s = O.get_image_blob(im, cfg)
        assert s == c[2] and blob.shape == ref.shape and blob.dtype == np.float32
        assert np.abs(blob - ref).max() <= BLOB_TOL, name


def test_image_blob_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for (h, w, s) in [(60, 100, 0.8), (75, 50, 1.6), (33, 50, 600 / 333.0), (20, 32, 3.0), (97, 131, 0.37)]:
        f = rng.randint(0, 256, (h, w, 3)).astype(np.float32)
        f -= O.PIXEL_MEANS
        ref = cv2.resize(f, None, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR)
        got = O.resize_linear_f32(f, s, s)
        assert got.shape == ref.shape and np.abs(got - ref).max() <= BLOB_TOL


def test_vgg16_conv5_shapes_ceil_mode():
    """Caffe pooling is ceil-mode (pooling_layer.cpp:93-95): a 75x125 input -> 5x8 conv5_3 map, post-ReLU."""
    from aznet_b200 import backbone
    w = backbone.make_vgg16_weights(seed=5, width_div=8)
    x = np.random.RandomState(1).standard_normal((1, 3, 75, 125)).astype(np.float32)
    y = O.vgg16_conv5(w, x)
    assert y.shape == (1, 64, 5, 8) and y.min() >= 0 and y.max() > 0
