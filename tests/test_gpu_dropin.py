"""GPU suite: the reference-facing Python entry points (aznet_b200.detect.test / utils) behave like
lib/detect/test.py and lib/utils/*.pyx: signatures, error behaviour, printed line, pickle formats, and
results against the oracle / golden vectors."""
import os
import pickle

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


@pytest.fixture()
def cfg():
    from aznet_b200.detect import config as C
    saved = (C.cfg.TEST.MAX_SIZE, C.cfg.SEAR.BATCH_SIZE, C.cfg.SEAR.FIXED_PROPOSAL_NUM, C.cfg.SEAR.APPEND_BOXES)
    C.cfg_set_path("pytest")
    C.cfg_set_mode("Test", 0.5)
    yield C.cfg
    C.cfg.TEST.MAX_SIZE, C.cfg.SEAR.BATCH_SIZE, C.cfg.SEAR.FIXED_PROPOSAL_NUM, C.cfg.SEAR.APPEND_BOXES = saved
    C.cfg_set_mode("Test", 0.5)


def test_cython_nms_dropin(dev, O):
    from aznet_b200.utils.cython_nms import nms
    d = synth.make_dets(1500, seed=4)
    assert nms(d, 0.5) == O.nms(d, 0.5)
    assert nms(d[::2], 0.3) == O.nms(np.ascontiguousarray(d[::2]), 0.3)       # strided view, like the typed buffer
    assert nms(np.zeros((0, 5), np.float32), 0.5) == []
    with pytest.raises(ValueError):
        nms(d.astype(np.float64), 0.5)
    with pytest.raises(ValueError):
        nms(d[:, 0], 0.5)
    with pytest.raises(TypeError):
        nms(d, np.float32(0.5))


def test_cython_div_dropin(dev, O):
    from aznet_b200.utils import cython_div
    b = synth.make_boxes(300, 480, 640, seed=12, lo=11, hi=300)
    assert np.array_equal(cython_div.divide_region(b, 10.0), O.divide_region(b, 10.0))
    assert np.array_equal(cython_div._sift_dup(b, 10.0), O.sift_dup(b, 10.0))
    assert cython_div.divide_region(np.zeros((0, 4)), 10.0).shape == (0, 4)
    with pytest.raises(ValueError):
        cython_div.divide_region(b.astype(np.float32), 10.0)


def test_bbox_pred_clip_dropin(dev, golden, cfg):
    from aznet_b200.detect import test as T
    g = golden["search"]
    # 1e-5 relative to the box scale: x1 = ctr - w/2 cancels, so the absolute floor is 1e-5 * O(100 px)
    np.testing.assert_allclose(T._bbox_pred(g["bbox_boxes"], g["bbox_deltas"]), g["bbox_pred"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(T._bbox_pred_clip(g["bbox_boxes"], g["bbox_deltas"], (600, 1000, 3)), g["bbox_clip"],
                               rtol=1e-5, atol=1e-3)
    a, c = T._unwrap_adj_pred(g["bbox_clip"].copy(), g["unwrap_scores_in"])
    assert np.array_equal(a, g["unwrap_boxes"]) and np.array_equal(c, g["unwrap_scores"])
    assert T._bbox_pred(np.zeros((0, 4)), np.zeros((0, 44), np.float32)).shape == (0, 44)


@pytest.mark.parametrize("name", ["d0_600x1000", "chunked_480x640", "tc_thresh_333x500", "append_375x500"])
def test_im_propose_host_route_matches_reference_golden(dev, golden, cfg, name, capsys):
    """Duck-typed foreign net (HashNet) -> host level loop with CUDA decode/divide; equals the reference's run."""
    from aznet_b200.detect import config as C
    from aznet_b200.detect import test as T
    g = golden["search"]
    H, W, max_size, bs, tz, rate, nprop, fixed = g[name + "_cfg"]
    cfg.TEST.MAX_SIZE, cfg.SEAR.BATCH_SIZE, cfg.SEAR.FIXED_PROPOSAL_NUM = int(max_size), int(bs), bool(fixed)
    cfg.SEAR.APPEND_BOXES = name.startswith("append")               # test.py:320-344, 403-406 (off by default)
    C.cfg_set_mode("Test", float(tz))
    if nprop > 0:
        cfg.SEAR.NUM_PROPOSALS = int(nprop)
    net = synth.HashNet(seed=11, zoom_rate=float(rate))
    Y = T.im_propose({"full": net, "fc": net}, np.zeros((int(H), int(W), 3), np.uint8))
    assert capsys.readouterr().out.strip().splitlines()[-1] == str(g[name + "_log"])
    ref = g[name + "_Y"]
    if not np.allclose(Y, ref, rtol=1e-5, atol=1e-5):
        np.testing.assert_allclose(Y[np.lexsort(Y.T[::-1])], ref[np.lexsort(ref.T[::-1])], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["train_375x500", "tz05_480x640", "stop_early_333x500"])
def test_tune_im_propose_host_route_matches_reference_golden(dev, golden, cfg, name, capsys):
    """detect.tune.im_propose with a foreign net = the reference's lib/detect/tune.py run."""
    from aznet_b200.detect import tune as U
    g = golden["tune"]
    H, W, max_size, bs, tz, rate, nprop = g[name + "_cfg"]
    cfg.TEST.MAX_SIZE, cfg.SEAR.BATCH_SIZE = int(max_size), int(bs)
    cfg.SEAR.Tz, cfg.SEAR.NUM_PROPOSALS = float(tz), int(nprop)
    net = synth.HashNet(seed=11, zoom_rate=float(rate))
    Y5, Bhis = U.im_propose({"full": net, "fc": net}, np.zeros((int(H), int(W), 3), np.uint8))
    assert capsys.readouterr().out.strip().splitlines()[-1] == str(g[name + "_log"])
    assert np.array_equal(Bhis.view(np.uint64), g[name + "_Bhis"].view(np.uint64))
    ref = g[name + "_Y5"]
    if not np.allclose(Y5, ref, rtol=1e-5, atol=1e-5):
        np.testing.assert_allclose(Y5[np.lexsort(Y5.T[::-1])], ref[np.lexsort(ref.T[::-1])], rtol=1e-5, atol=1e-5)


def test_tune_thresh_dropin(dev, golden, cfg, tmp_path, capsys):
    """tune_thresh: (a) foreign net over the golden image set -> the reference's thresh.pkl value;
    (b) device route (Net) = host route over the same nets; thresh.pkl round-trips through cfg_load_thresh."""
    import cv2
    from aznet_b200.detect import config as C
    from aznet_b200.detect import tune as U
    g = golden["tune"]
    paths = []
    for i, (h, w) in enumerate(g["thresh_shapes"]):
        paths.append(str(tmp_path / ("%d.png" % i)))
        cv2.imwrite(paths[-1], synth.make_images(1, int(h), int(w), seed=40 + i)[0])

    class Imdb:
        name = "synth_tune"
        image_index = list(range(len(paths)))

        def image_path_at(self, i):
            return paths[i]
    saved_root, saved_per = cfg.ROOT_DIR, cfg.TRAIN.ANCHORS_PER_IMG
    cfg.ROOT_DIR = str(tmp_path)
    try:
        C.cfg_set_mode("Train")
        hnet = synth.HashNet(seed=11, zoom_rate=0.5)
        th = U.tune_thresh({"full": hnet, "fc": hnet}, Imdb())
        assert float(th) == float(g["thresh_per20"])
        assert float(C.cfg_load_thresh(os.path.join(C.get_output_dir(Imdb(), hnet), "thresh.pkl"))) == float(th)
        cfg.TRAIN.ANCHORS_PER_IMG = 1000000
        assert U.tune_thresh({"full": hnet, "fc": hnet}, Imdb()) == -np.inf
        cfg.TRAIN.ANCHORS_PER_IMG = 5
        az, _, _, _ = _small_nets(dev)
        th_fast = U.tune_thresh(az, Imdb())

        class Foreign(dict):
            pass
        wrap = lambda n: type("W", (), {"forward": n.forward, "blobs": n.blobs, "name": n.name})()
        th_host = U.tune_thresh(Foreign(full=wrap(az["full"]), fc=wrap(az["fc"])), Imdb())
        assert np.isfinite(th_fast) and abs(float(th_fast) - float(th_host)) < 2e-3
        out = capsys.readouterr().out
        assert "the threshold is set to" in out and "im_tune: 6/6" in out
    finally:
        cfg.ROOT_DIR, cfg.TRAIN.ANCHORS_PER_IMG = saved_root, saved_per


def _small_nets(dev, num_classes=6):
    from aznet_b200 import backbone, net
    bw = backbone.make_vgg16_weights(seed=5, width_div=8)                 # conv5_3 has 64 channels
    bb = backbone.VGG16Backbone(bw, dev)
    azw = synth.make_az_weights(seed=3, C=64, h6=256, h71=96, h72=32, zoom_bias=0.0)
    frw = synth.make_frcnn_weights(seed=4, num_classes=num_classes, C=64, h6=256, h7=128)
    frw["cls_score"] = (frw["cls_score"][0] * np.float32(0.02), frw["cls_score"][1])      # O(1) logits: unsaturated softmax
    frw["bbox_pred"] = (frw["bbox_pred"][0] * np.float32(0.05), frw["bbox_pred"][1])      # |deltas| << 1: boxes stay box-sized
    az = {"full": net.Net(azw, "az", backbone=bb, name="az_small"), "fc": net.Net(azw, "az", name="az_small")}
    fr = {"full": net.Net(frw, "frcnn", backbone=bb, name="frcnn_small"), "fc": net.Net(frw, "frcnn", name="frcnn_small")}
    return az, fr, azw, frw


def test_im_propose_fast_route_equals_host_route(dev, cfg, capsys, monkeypatch):
    """Device-resident engine vs one-forward-per-level host loop over the same Net objects."""
    from aznet_b200.detect import test as T
    _same_blob_for_both_routes(monkeypatch, dev)
    az, _, _, _ = _small_nets(dev)
    im = synth.make_images(1, 240, 320, seed=3)[0]
    Y_fast, conv = T.im_propose(az, im, return_conv=True)
    line_fast = capsys.readouterr().out.strip().splitlines()[-1]
    assert conv["conv5_3"].shape[:2] == (1, 64) and conv["conv5_3"].dtype == np.float32

    class Foreign(dict):                      # hides the Net type -> forces the host route
        pass
    wrap = lambda n: type("W", (), {"forward": n.forward, "blobs": n.blobs, "name": n.name})()
    Y_host = T.im_propose(Foreign(full=wrap(az["full"]), fc=wrap(az["fc"])), im)
    line_host = capsys.readouterr().out.strip().splitlines()[-1]
    assert line_fast == line_host
    np.testing.assert_allclose(Y_fast, Y_host, rtol=1e-5, atol=1e-4)


def test_net_forward_contract(dev):
    az, fr, _, _ = _small_nets(dev)
    fc = az["fc"]
    conv = synth.make_conv_maps(1, 64, 15, 20, seed=2)
    rois = synth.make_rois(9, 240, 320, seed=1)
    with pytest.raises(Exception, match="do not match net inputs"):
        fc.forward(rois=rois)
    fc.blobs["rois"].reshape(3, 5)
    fc.blobs["conv5_3"].reshape(*conv.shape)
    with pytest.raises(Exception, match="not batch sized"):
        fc.forward(rois=rois, conv5_3=conv)
    fc.blobs["rois"].reshape(*rois.shape)
    out = fc.forward(rois=rois, conv5_3=conv, blobs=["conv5_3"])
    assert out["zoom_prob"].shape == (9, 1) and out["adj_prob"].shape == (9, 11) and out["adj_bbox"].shape == (9, 44)
    assert out["conv5_3"] is conv and out["adj_prob"].dtype == np.float32
    f2 = fr["fc"]
    f2.blobs["rois"].reshape(*rois.shape)
    f2.blobs["conv5_3"].reshape(*conv.shape)
    o2 = f2.forward(rois=rois, conv5_3=conv)
    assert o2["cls_prob"].shape == (9, 6) and o2["bbox_pred"].shape == (9, 24)
    np.testing.assert_allclose(o2["cls_prob"].sum(1), 1.0, atol=1e-5)


def test_im_detect_vs_oracle(dev, O, cfg):
    from aznet_b200.detect import test as T
    _, fr, _, frw = _small_nets(dev)
    im = synth.make_images(1, 240, 320, seed=5)[0]
    boxes = synth.make_boxes(120, 240, 320, seed=6, lo=12, hi=200)
    scores, pred = T.im_detect(fr, im, boxes, 6)
    assert scores.shape == (120, 6) and pred.shape == (120, 24) and pred.dtype == np.float64
    # oracle on the same conv map (taken from our backbone) with bf16-rounded weights, fp32 math
    data, _ = T._get_image_blob(im)
    conv = fr["full"].conv_from_data(data)[0].cpu().numpy()
    bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    wq = {k: (bf(v[0]), v[1]) for k, v in frw.items()}
    ocfg = O.OracleCfg()
    # (1) the oracle with the product's bf16 activation storage emulated: only the fp32 summation order differs
    onet = O.OracleNet(wq, "frcnn", cfg=ocfg, act_round=O.round_bf16)
    s_ref, p_ref, _ = O.frcnn_forward({"full": onet, "fc": onet}, im.shape, boxes, 6, {"conv5_3": bf(conv)}, ocfg)
    np.testing.assert_allclose(scores, s_ref, atol=2e-2)           # a flipped bf16 rounding of one activation moves a probability by ~1e-2
    np.testing.assert_allclose(pred, p_ref, rtol=5e-3, atol=0.5)
    # (2) the reference's fp32 blobs: stated tolerance for bf16 activations.  The logits here are O(10) (the conv
    # map of a random-init backbone is not normalised), so a 2^-9 relative activation error moves a probability
    # by up to ~3e-2
    onet = O.OracleNet(wq, "frcnn", cfg=ocfg)
    s_ref, p_ref, _ = O.frcnn_forward({"full": onet, "fc": onet}, im.shape, boxes, 6, {"conv5_3": bf(conv)}, ocfg)
    np.testing.assert_allclose(scores, s_ref, atol=4e-2)
    assert np.mean(np.abs(scores - s_ref)) < 4e-3
    np.testing.assert_allclose(pred, p_ref, rtol=2e-2, atol=2.0)


def test_drivers_and_pickle_formats(dev, O, cfg, tmp_path, capsys):
    """test_proposals -> proposals.pkl {'boxes','time','recall'}; test_net -> detections.pkl (list[cls][img] of
    float32 [n,5]); apply_nms result equals the oracle's apply_nms on the same detections, bit for bit."""
    import cv2
    from aznet_b200.detect import config as C
    from aznet_b200.detect import test as T
    az, fr, _, _ = _small_nets(dev)
    paths = []
    for i, im in enumerate(synth.make_images(3, 200, 300, seed=40)):
        p = str(tmp_path / ("im%d.png" % i))
        cv2.imwrite(p, im)
        paths.append(p)
    imdb = synth.SyntheticImdb(paths, num_classes=6)
    C.cfg.ROOT_DIR = str(tmp_path)
    T.test_proposals(az, imdb)
    out = capsys.readouterr().out
    assert "im_prop: 3/3" in out and "The average proposal generation time is" in out
    prop_file = os.path.join(C.get_output_dir(imdb, az["full"]), "proposals.pkl")
    prop = pickle.load(open(prop_file, "rb"))
    assert set(prop.keys()) == {"boxes", "time", "recall"} and len(prop["boxes"]) == 3
    assert all(b.dtype == np.float64 and b.shape[1] == 4 and 0 < b.shape[0] <= 300 for b in prop["boxes"])
    T.test_net(fr, prop_file, imdb)
    out = capsys.readouterr().out
    assert "Applying NMS to all detections" in out and "Evaluating detections" in out
    dets = pickle.load(open(os.path.join(C.get_output_dir(imdb, fr["full"]), "detections.pkl"), "rb"))
    assert len(dets) == 6 and len(dets[0]) == 3 and dets[0][0] == []
    assert all(d.dtype == np.float32 and d.shape[1] == 5 for c in dets[1:] for d in c)
    nms_dets, out_dir = imdb.evaluated
    ref = O.apply_nms(dets, C.cfg.TEST.NMS)
    for c in range(6):
        for i in range(3):
            a, b = nms_dets[c][i], ref[c][i]
            assert (isinstance(a, list) and isinstance(b, list)) or np.array_equal(a, b), (c, i)
    T.test_net_shared(az, fr, imdb)
    assert "The average detection time is" in capsys.readouterr().out


def test_test_net_device_route_equals_host_route(dev, cfg, tmp_path, capsys, monkeypatch):
    """test_net / test_net_shared with Net detectors (DetectEngine + set-wide finish on the device) against the
    reference's host loop over the same Net objects (numpy selection, heap thresholds, per-class NMS calls)."""
    import cv2
    from aznet_b200.detect import config as C
    from aznet_b200.detect import test as T
    _same_blob_for_both_routes(monkeypatch, dev)
    az, fr, _, _ = _small_nets(dev)
    paths = []
    for i, im in enumerate(synth.make_images(4, 200, 300, seed=60)):
        p = str(tmp_path / ("im%d.png" % i))
        cv2.imwrite(p, im)
        paths.append(p)
    C.cfg.ROOT_DIR = str(tmp_path)
    imdb = synth.SyntheticImdb(paths, num_classes=6)
    T.test_proposals(az, imdb)
    prop_file = os.path.join(C.get_output_dir(imdb, az["full"]), "proposals.pkl")
    det_file = os.path.join(C.get_output_dir(imdb, fr["full"]), "detections.pkl")

    class Foreign(dict):                      # hides the Net type -> forces the host route
        pass
    wrap = lambda n: type("W", (), {"forward": n.forward, "blobs": n.blobs, "name": n.name})()

    def run(nets, shared):
        if shared:
            T.test_net_shared(az, nets, imdb)
        else:
            T.test_net(nets, prop_file, imdb)
        capsys.readouterr()
        return pickle.load(open(det_file, "rb")), imdb.evaluated[0]

    match = RowMatcher(atol=0.5)
    for shared in (False, True):
        pre_d, nms_d = run(fr, shared)
        pre_h, nms_h = run(Foreign(full=wrap(fr["full"]), fc=wrap(fr["fc"])), shared)
        for a_set, b_set in ((pre_d, pre_h), (nms_d, nms_h)):
            for j in range(1, 6):
                for i in range(4):
                    a, b = a_set[j][i], b_set[j][i]
                    match.add(a, b)          # the batched route runs 4 images per GEMM, the host loop one
    match.check(0.95, min_total=100)


def test_batched_drivers_mixed_shapes_and_in_memory_imdb(dev, cfg, tmp_path, capsys, monkeypatch):
    """test_proposals / test_net / test_net_shared on a database of TWO image shapes (7 images -> two same-shape
    batches, padded to 8 and 2): results land at their image indices, equal the per-image im_propose / im_detect of the
    same nets up to bf16 rounding flips, the reference's lines are printed once per image, and an imdb that offers
    image_at(i) (in-memory arrays) gives the same proposals as the PNG files."""
    import cv2
    from aznet_b200.detect import config as C
    from aznet_b200.detect import test as T
    _same_blob_for_both_routes(monkeypatch, dev)
    az, fr, _, _ = _small_nets(dev)
    ims = synth.make_images(5, 200, 300, seed=80) + synth.make_images(2, 240, 200, seed=90)
    order = [0, 5, 1, 2, 6, 3, 4]                              # shapes interleaved in database order
    ims = [ims[k] for k in order]
    paths = []
    for i, im in enumerate(ims):
        p = str(tmp_path / ("m%d.png" % i))
        cv2.imwrite(p, im)
        paths.append(p)
    C.cfg.ROOT_DIR = str(tmp_path)
    imdb = synth.SyntheticImdb(paths, num_classes=6, name="mixed")
    T.test_proposals(az, imdb)
    out = capsys.readouterr().out
    assert T.test_proposals.last_stats["route"] == "batched" and T.test_proposals.last_stats["batches"] == 2
    assert out.count("proposals, evaluate") == 7 and "im_prop: 7/7" in out and "im_prop: 1/7" in out
    prop_file = os.path.join(C.get_output_dir(imdb, az["full"]), "proposals.pkl")
    prop = pickle.load(open(prop_file, "rb"))
    match = RowMatcher(atol=0.05)
    for i, im in enumerate(ims):
        assert prop["boxes"][i].dtype == np.float64 and 0 < prop["boxes"][i].shape[0] <= 300
        match.add(prop["boxes"][i], T.im_propose(az, im))     # single-image engine on the same image
    match.check(0.97, min_total=200)
    capsys.readouterr()
    mem = synth.InMemoryImdb(ims, num_classes=6, name="mixed_mem")
    T.test_proposals(az, mem)
    capsys.readouterr()
    prop_mem = pickle.load(open(os.path.join(C.get_output_dir(mem, az["full"]), "proposals.pkl"), "rb"))
    for i in range(7):
        assert np.array_equal(prop_mem["boxes"][i], prop["boxes"][i])        # same batches, same kernels: bit-identical
    # detection over the saved proposals; one image without proposals is skipped like the reference does
    prop["boxes"][3] = np.zeros((0, 4))
    pickle.dump(prop, open(prop_file, "wb"))
    T.test_net(fr, prop_file, imdb)
    out = capsys.readouterr().out
    assert T.test_net.last_stats["route"] == "batched" and "im_detect: 6/7" in out and "im_detect: 7/7" not in out
    dets = pickle.load(open(os.path.join(C.get_output_dir(imdb, fr["full"]), "detections.pkl"), "rb"))
    assert all(isinstance(dets[j][3], list) and dets[j][3] == [] for j in range(6))
    match = RowMatcher(atol=0.5)
    for i in (0, 1, 4):
        s_one, p_one = T.im_detect(fr, ims[i], prop["boxes"][i], 6)
        for j in range(1, 6):
            top = np.argsort(-s_one[:, j], kind="stable")[:100]
            one = np.hstack((p_one[top, 4 * j:4 * j + 4], s_one[top, j:j + 1]))
            thr = dets[j][i][:, 4].min() if len(dets[j][i]) else np.inf
            match.add(dets[j][i], one[one[:, 4] >= thr - 1e-3])
    match.check(0.95, min_total=200)
    T.test_net_shared(az, fr, imdb)
    out = capsys.readouterr().out
    assert T.test_net_shared.last_stats["route"] == "batched" and "im_detect: 7/7" in out
    assert out.count("proposals, evaluate") == 7


def test_test_net_without_cfg_set_mode(dev, tmp_path, capsys):
    """tools/test_det_net.py calls test_net(nets, prop_file, imdb) WITHOUT cfg_set_mode, so SEAR.NUM_PROPOSALS and
    SEAR.Tz do not exist (lib/detect/config.py:272-280); the reference's test_net never reads them."""
    import cv2
    from aznet_b200.detect import config as C
    from aznet_b200.detect import test as T
    _, fr, _, _ = _small_nets(dev)
    saved = {k: C.cfg.SEAR.pop(k) for k in ("NUM_PROPOSALS", "Tz") if k in C.cfg.SEAR}
    try:
        assert not hasattr(C.cfg.SEAR, "NUM_PROPOSALS")
        paths = []
        for i, im in enumerate(synth.make_images(2, 200, 300, seed=95)):
            paths.append(str(tmp_path / ("n%d.png" % i)))
            cv2.imwrite(paths[-1], im)
        C.cfg.ROOT_DIR = str(tmp_path)
        C.cfg_set_path("pytest_nomode")
        imdb = synth.SyntheticImdb(paths, num_classes=6, name="nomode")
        prop_file = str(tmp_path / "proposals.pkl")
        boxes = [synth.make_boxes(350, 200, 300, seed=96 + i, lo=12, hi=150) for i in range(2)]   # longer than 300 rows
        pickle.dump({"boxes": boxes, "time": 0.0, "recall": 0}, open(prop_file, "wb"))
        T.test_net(fr, prop_file, imdb)
        assert "Evaluating detections" in capsys.readouterr().out and imdb.evaluated is not None
    finally:
        for k, v in saved.items():
            C.cfg.SEAR[k] = v
