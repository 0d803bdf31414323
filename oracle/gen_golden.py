"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Run in the authoring container only.

Generates tests/golden/*.npz by executing the REFERENCE'S OWN CODE on seeded inputs:
  * lib/utils/div.pyx / nms.pyx compiled unmodified (oracle/build_ref.py),
  * lib/detect/test.py converted py2->py3 in oracle/_ref/pyref (never committed), driven with
    aznet_b200.synth.HashNet (integer-hash net outputs, bit-reproducible on any machine).
The fixtures store inputs' seeds + the reference's outputs; tests/ compare the oracle
restatement (CPU suite) and the CUDA path (gpu suite) against them.

    python oracle/gen_golden.py                 # rewrites tests/golden/
    python oracle/gen_golden.py --only blob     # one fixture file (div | nms | search | blob | tune | detect | layers)
"""
from __future__ import annotations

import io
import os
import sys
import contextlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from aznet_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def load_reference():
    """Build oracle/_ref from /root/reference and import the converted lib/detect/test.py with the compiled Cython
    modules behind it (oracle/build_ref.py::load_pyref)."""
    assert build_ref.build(), "/root/reference is required to generate goldens"
    return build_ref.load_pyref()


def gen_div(div):
    rng = np.random.default_rng(21)
    cases = {}
    cases["root_600x1000"] = np.array([[0, 0, 999, 599.]])
    cases["root_375x500"] = np.array([[0, 0, 499, 374.]])
    cases["square"] = np.array([[10, 20, 109, 119.]])
    cases["tall"] = np.array([[5, 5, 24, 204.]])
    b = synth.make_boxes(200, 600, 1000, seed=5, lo=12, hi=500)
    cases["random200"] = b
    cases["dups"] = np.vstack([b[:20], b[:20] + 0.4, b[:20]])
    # full-zoom cascade from the 600x1000 root: levels 1..5 (sizes 1, 8, 32, 134, 564)
    lv = cases["root_600x1000"]
    out = {}
    for name, r in cases.items():
        out["in_" + name] = r
        out["out_" + name] = div.divide_region(np.ascontiguousarray(r, dtype=np.float64), 10.0)
    sizes = [1]
    for k in range(4):
        lv = div.divide_region(lv, 10.0)
        sizes.append(lv.shape[0])
        out["cascade_%d" % (k + 2)] = lv
    out["cascade_sizes"] = np.array(sizes)
    out["sift_in"] = np.round(rng.uniform(0, 300, (500, 4)), 1)
    out["sift_out"] = div._sift_dup(out["sift_in"], 10.0)
    np.savez_compressed(os.path.join(GOLD, "div.npz"), **out)
    print("div: cascade sizes", sizes)


def gen_nms(nms):
    out = {}
    for n in (1, 2, 17, 300, 2000):
        dets = synth.make_dets(n, seed=3)
        for th in (0.3, 0.5, 0.7):
            keep = nms.nms(dets, th)
            out["keep_n%d_t%d" % (n, int(th * 10))] = np.array(keep, dtype=np.int64)
    # boxes engineered so that some IoUs land exactly on the threshold (suppress iff ovr >= thresh)
    d = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 4, 0.8], [0, 5, 9, 9, 0.7], [20, 20, 29, 29, 0.6],
                  [20, 20, 29, 24, 0.5]], dtype=np.float32)
    out["edge_dets"] = d
    out["edge_keep_t5"] = np.array(nms.nms(d, 0.5), dtype=np.int64)
    out["edge_keep_t51"] = np.array(nms.nms(d, 0.51), dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "nms.npz"), **out)
    print("nms: n2000 keeps", [len(out["keep_n2000_t%d" % t]) for t in (3, 5, 7)])


def gen_search(rtest, rconfig):
    """Reference im_propose (lib/detect/test.py:346-414) + HashNet on several image shapes/configs."""
    cfg = rconfig.cfg
    out = {}
    cases = [
        # name, (H, W), MAX_SIZE, BATCH_SIZE, Tz, zoom_rate, num_proposals, fixed
        ("d0_600x1000", (600, 1000), 1000, 10000, 0.5, 0.5, 300, True),
        ("voc_600x1000", (600, 1000), 800, 1000, 0.5, 0.4, 300, True),
        ("fullzoom_600x1000", (600, 1000), 1000, 10000, 0.0, 0.5, 2000, True),
        ("small_375x500", (375, 500), 1000, 10000, 0.5, 0.6, 300, True),
        ("chunked_480x640", (480, 640), 800, 50, 0.0, 0.5, 300, True),
        ("tc_thresh_333x500", (333, 500), 1000, 10000, 0.5, 0.5, None, False),
        ("nozoom_600x1000", (600, 1000), 1000, 10000, 0.999, 0.0, 300, True),
        ("append_375x500", (375, 500), 1000, 10000, 0.5, 0.5, 300, True),       # SEAR.APPEND_BOXES (test.py:320-344, 403-406)
    ]
    for name, shape, max_size, bs, tz, rate, nprop, fixed in cases:
        cfg.SEAR.APPEND_BOXES = name.startswith("append")
        cfg.TEST.MAX_SIZE = max_size
        cfg.SEAR.BATCH_SIZE = bs
        cfg.SEAR.FIXED_PROPOSAL_NUM = fixed
        rconfig.cfg_set_mode("Test", tz)
        if nprop is not None:
            cfg.SEAR.NUM_PROPOSALS = nprop
        net = synth.HashNet(seed=11, zoom_rate=rate)
        im = np.zeros(shape + (3,), dtype=np.uint8)
        conv = {"conv5_3": np.zeros((1, 1, 2, 2), np.float32)}
        # the reference computes `conv` at level 1 through net['full']; HashNet ignores the image, so
        # 'full' and 'fc' are the same object and the image blob path (cv2.resize) still runs.
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            Y = rtest.im_propose({"full": net, "fc": net}, im)
        line = buf.getvalue().strip()
        out[name + "_Y"] = Y
        out[name + "_log"] = np.array(line)
        out[name + "_cfg"] = np.array([shape[0], shape[1], max_size, bs, tz, rate, -1 if nprop is None else nprop,
                                       int(fixed)], dtype=np.float64)
        print("search:", name, line)
    cfg.SEAR.APPEND_BOXES = False
    # component vectors
    rng = np.random.default_rng(9)
    boxes = synth.make_boxes(64, 600, 1000, seed=8)
    deltas = (rng.standard_normal((64, 44)) * 0.3).astype(np.float32)
    pred = rtest._bbox_pred(boxes, deltas)
    out["bbox_boxes"], out["bbox_deltas"], out["bbox_pred"] = boxes, deltas, pred.copy()
    out["bbox_clip"] = rtest._clip_boxes(pred.copy(), (600, 1000, 3))
    scores = rng.uniform(0, 1, (64, 11)).astype(np.float32)
    a, c = rtest._unwrap_adj_pred(out["bbox_clip"], scores)
    out["unwrap_scores_in"], out["unwrap_boxes"], out["unwrap_scores"] = scores, a, c
    np.savez_compressed(os.path.join(GOLD, "search.npz"), **out)


BLOB_CASES = [  # name, image shape, TEST.SCALES, TEST.MAX_SIZE
    ("up_1p6", (30, 50), (48,), 1000),
    ("capped_by_max_size", (45, 75), (60,), 80),
    ("down_portrait", (64, 48), (40,), 1000),
    ("odd_ratio", (37, 53), (61,), 1000),
]


def gen_blob(rtest, rconfig):
    """The reference's own _get_image_blob (lib/detect/test.py:27-59: mean subtraction + cv2.resize INTER_LINEAR +
    im_list_to_blob) on seeded uint8 images -> tests/golden/blob.npz."""
    cfg = rconfig.cfg
    out = {}
    keep = (cfg.TEST.SCALES, cfg.TEST.MAX_SIZE)
    for k, (name, shape, scales, max_size) in enumerate(BLOB_CASES):
        cfg.TEST.SCALES, cfg.TEST.MAX_SIZE = scales, max_size
        im = np.random.RandomState(1000 + k).randint(0, 256, shape + (3,)).astype(np.uint8)
        blob, factors = rtest._get_image_blob(im)
        out[name + "_im"], out[name + "_blob"] = im, blob.astype(np.float32)
        out[name + "_cfg"] = np.array([scales[0], max_size, float(factors[0])], dtype=np.float64)
        print("blob:", name, im.shape, "->", blob.shape, "scale", float(factors[0]))
    cfg.TEST.SCALES, cfg.TEST.MAX_SIZE = keep
    np.savez_compressed(os.path.join(GOLD, "blob.npz"), **out)


TUNE_CASES = [  # name, (H, W), MAX_SIZE, BATCH_SIZE, Tz, zoom_rate, NUM_PROPOSALS
    ("train_375x500", (375, 500), 1000, 10000, 0.0, 0.5, 2000),      # tools/set_thresh.py: cfg_set_mode('Train') -> Tz = 0
    ("train_voc_600x1000", (600, 1000), 800, 1000, 0.0, 0.4, 2000),
    ("tz05_480x640", (480, 640), 1000, 10000, 0.5, 0.5, 300),        # tools/diagnose_prop.py: Test mode
    ("stop_early_333x500", (333, 500), 1000, 10000, 0.999, 0.0, 300),
]


def gen_tune(rconfig):
    """The reference's diagnostic im_propose + tune_thresh (lib/detect/tune.py:256-366) with HashNet."""
    import tempfile
    import pickle
    import cv2
    import detect.tune as rtune
    cfg = rconfig.cfg
    out = {}
    for name, shape, max_size, bs, tz, rate, nprop in TUNE_CASES:
        cfg.TEST.MAX_SIZE = max_size
        cfg.SEAR.BATCH_SIZE = bs
        cfg.SEAR.Tz = tz
        cfg.SEAR.NUM_PROPOSALS = nprop
        net = synth.HashNet(seed=11, zoom_rate=rate)
        im = np.zeros(shape + (3,), dtype=np.uint8)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            Y5, Bhis = rtune.im_propose({"full": net, "fc": net}, im)
        out[name + "_Y5"], out[name + "_Bhis"] = Y5, Bhis
        out[name + "_log"] = np.array(buf.getvalue().strip())
        out[name + "_cfg"] = np.array([shape[0], shape[1], max_size, bs, tz, rate, nprop], dtype=np.float64)
        print("tune:", name, buf.getvalue().strip(), "history", Bhis.shape)
    # tune_thresh over a small image set (three sizes, written as PNGs for the reference's cv2.imread)
    cfg.TEST.MAX_SIZE, cfg.SEAR.BATCH_SIZE = 1000, 10000
    rconfig.cfg_set_mode("Train")
    shapes = [(375, 500), (333, 500), (375, 500), (480, 640), (375, 500), (333, 500)]
    for per_img in (20, 1000000):
        cfg.TRAIN.ANCHORS_PER_IMG = per_img
        with tempfile.TemporaryDirectory() as tmp:
            paths = []
            for i, sh in enumerate(shapes):
                paths.append(os.path.join(tmp, "%d.png" % i))
                cv2.imwrite(paths[-1], np.zeros(sh + (3,), np.uint8))

            class Imdb:
                name = "synth"
                image_index = list(range(len(shapes)))
                roidb = None

                def image_path_at(self, i):
                    return paths[i]

                def gt_roidb(self):
                    return None
            cfg.ROOT_DIR = tmp
            rconfig.cfg_set_path("golden")
            net = synth.HashNet(seed=11, zoom_rate=0.5)
            with contextlib.redirect_stdout(io.StringIO()):
                rtune.tune_thresh({"full": net, "fc": net}, Imdb())
            with open(os.path.join(rconfig.get_output_dir(Imdb(), net), "thresh.pkl"), "rb") as f:
                th = pickle.load(f)
        out["thresh_per%d" % per_img] = np.array(float(th))
        print("tune_thresh: ANCHORS_PER_IMG", per_img, "->", th)
    cfg.TRAIN.ANCHORS_PER_IMG = 20
    out["thresh_shapes"] = np.array(shapes)
    np.savez_compressed(os.path.join(GOLD, "tune.npz"), **out)


# name, num_classes, [(H, W)] per image, MAX_SIZE, BATCH_SIZE, proposals per image, NMS threshold
DETECT_CASES = [
    ("voc_3sizes", 21, [(600, 1000), (375, 500), (480, 640), (600, 1000), (375, 500), (480, 640)], 1000, 10000,
     [300, 257, 300, 150, 300, 64], 0.3),          # six PNGs of three sizes; max_per_set = 40 * 6 < pushed -> thresholds
    ("coco_480x640", 81, [(480, 640)] * 4, 800, 10000, [300, 300, 12, 0], 0.5),   # 81 classes; an image w/o proposals
    ("voc_chunked_375x500", 21, [(375, 500)] * 2, 1000, 100, [300, 150], 0.3),    # dedup per BATCH_SIZE = 100 boxes
]


def detect_case_proposals(shapes, counts, seed=21):
    """Per-image proposal arrays of a DETECT_CASES row (same generator for the goldens and the tests): image i uses
    synth.make_detect_proposals with seed + i on its own shape."""
    return [synth.make_detect_proposals([c], h, w, seed=seed + i)[0, :c] for i, ((h, w), c) in enumerate(zip(shapes, counts))]


def gen_detect(rtest, rconfig):
    """The reference's own detection path with HashDetNet: im_detect -> _frcnn_forward (lib/detect/test.py:259-318,
    416-430), test_net's per-class selection + heap + final filter (:541-668, detections.pkl) and apply_nms
    (:467-484, what imdb.evaluate_detections receives)."""
    import pickle
    import tempfile
    import cv2
    cfg = rconfig.cfg
    out = {}
    keep = (cfg.TEST.MAX_SIZE, cfg.SEAR.BATCH_SIZE, cfg.TEST.NMS)
    for name, ncls, shapes, max_size, bs, counts, nms_t in DETECT_CASES:
        cfg.TEST.MAX_SIZE, cfg.SEAR.BATCH_SIZE, cfg.TEST.NMS = max_size, bs, nms_t
        props = detect_case_proposals(shapes, counts)
        net = synth.HashDetNet(seed=13, num_classes=ncls)
        with tempfile.TemporaryDirectory() as tmp:
            paths = []
            for i, sh in enumerate(shapes):
                paths.append(os.path.join(tmp, "%d.png" % i))
                cv2.imwrite(paths[-1], np.zeros(sh + (3,), np.uint8))
            seen = {}

            class Imdb:
                name = "synth"
                image_index = list(range(len(shapes)))
                num_classes = ncls
                classes = ["c%d" % j for j in range(ncls)]

                def image_path_at(self, i):
                    return paths[i]

                def evaluate_detections(self, all_boxes, output_dir):
                    seen["nms"] = all_boxes

            cfg.ROOT_DIR = tmp
            rconfig.cfg_set_path("golden")
            prop_file = os.path.join(tmp, "proposals.pkl")
            with open(prop_file, "wb") as f:
                pickle.dump({"boxes": props, "time": 0.0, "recall": 0}, f)
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                rtest.test_net({"full": net}, prop_file, Imdb())
            with open(os.path.join(rconfig.get_output_dir(Imdb(), net), "detections.pkl"), "rb") as f:
                dets = pickle.load(f)
            # im_detect of the smallest non-empty image (bounded fixture size)
            i0 = min((c, i) for i, c in enumerate(counts) if c > 0)[1]
            sc, pb = rtest.im_detect({"full": net}, cv2.imread(paths[i0]), props[i0], ncls)
        out[name + "_imdet_index"] = np.array(i0)
        out[name + "_imdet_scores"] = sc.astype(np.float32)        # f32 values promoted to f64 by vstack (:266, 315)
        assert np.array_equal(out[name + "_imdet_scores"].astype(np.float64), sc)
        out[name + "_imdet_boxes"] = pb
        for tag, ab in (("det", dets), ("nms", seen["nms"])):
            rows, cnt = [], np.zeros((ncls, len(shapes)), np.int32)
            for j in range(ncls):
                for i in range(len(shapes)):
                    d = ab[j][i]
                    if isinstance(d, list):
                        continue
                    cnt[j, i] = d.shape[0]
                    assert d.dtype == np.float32
                    rows.append(d)
            out["%s_%s_rows" % (name, tag)] = np.vstack(rows) if rows else np.zeros((0, 5), np.float32)
            out["%s_%s_count" % (name, tag)] = cnt
        out[name + "_last_line"] = np.array(buf.getvalue().strip().split("\n")[-1])
        print("detect:", name, "rows", out[name + "_det_rows"].shape[0], "after nms", out[name + "_nms_rows"].shape[0],
              "|", buf.getvalue().strip().split("\n")[-1])
    cfg.TEST.MAX_SIZE, cfg.SEAR.BATCH_SIZE, cfg.TEST.NMS = keep
    np.savez_compressed(os.path.join(GOLD, "detect.npz"), **out)


def layer_roi_cases():
    """ROI sets of the Caffe-layer goldens: edge cases (coordinates landing on .5 after the 1/16 scale, out-of-image,
    malformed, whole image) and the 739 'natural' regions of the full-zoom cascade of a 600x1000 image."""
    from oracle import az_oracle as O
    edge = synth.make_rois(200, 600, 1000, seed=5, n_img=2)
    edge[:10, 1:] = np.round(edge[:10, 1:] / 8) * 8
    edge[10:14, 1:] += 900
    edge[14] = [0, 50, 50, 40, 40]
    edge[15] = [1, -300, -200, -20, -30]
    edge[16] = [0, 0, 0, 999, 599]
    edge[17] = [1, 0, 0, 1000, 600]
    lv, regs = np.array([[0, 0, 999, 599.]]), []
    for _ in range(5):
        regs.append(lv)
        lv = O.divide_region(lv, 10.0)
    nat = np.vstack(regs)
    nat = np.hstack([np.zeros((nat.shape[0], 1)), nat]).astype(np.float32)
    return edge, nat


def gen_layers():
    """Forward_cpu of the reference's own layer sources (oracle/_ref/libcaffe_layers_ref.so, oracle/ref_caffe.py)."""
    from oracle import ref_caffe as RC
    out = {}
    feat = synth.make_conv_maps(2, 8, 38, 63, seed=7)
    feat[1] -= 0.5                                          # negative values as well (pre-ReLU style maps)
    feat[1, 0, 3, 4] = np.nan
    feat[1, 1, 10:20, 10:30] = -0.0
    edge, nat = layer_roi_cases()
    for tag, rois in (("edge", edge), ("natural", nat)):
        o, am = RC.roi_pool_fwd(feat, rois, want_argmax=True)
        out["roi_%s_rois" % tag], out["roi_%s_out" % tag], out["roi_%s_argmax" % tag] = rois, o, am.astype(np.int16 if am.max() < 32768 else np.int32)
    out["roi_feat_seed"] = np.array(7)
    x = np.random.default_rng(1).standard_normal((3, 40, 7, 7)).astype(np.float32)
    x[0, :, 0, 0] = 0
    out["grn_in"] = x
    with np.errstate(all="ignore"):
        out["grn_out"] = RC.grn(x)
    z = np.linspace(-30, 30, 2001).astype(np.float32)
    out["sigmoid_in"], out["sigmoid_out"] = z, RC.sigmoid(z)
    out["relu_out"] = RC.relu(z)
    s = (np.random.default_rng(2).standard_normal((60, 81)) * 5).astype(np.float32)
    out["softmax_in"], out["softmax_out"] = s, RC.softmax(s)
    xx = np.random.default_rng(3).standard_normal((9, 4, 7, 7)).astype(np.float32)
    w = np.random.default_rng(4).standard_normal((13, 196)).astype(np.float32)
    b = np.random.default_rng(5).standard_normal(13).astype(np.float32)
    out["ip_x"], out["ip_w"], out["ip_b"], out["ip_out"] = xx, w, b, RC.inner_product(xx, w, b)
    pm = np.random.default_rng(6).standard_normal((1, 3, 75, 125)).astype(np.float32)
    out["maxpool_in"], out["maxpool_out"] = pm, RC.max_pool(pm)
    np.savez_compressed(os.path.join(GOLD, "caffe_layers.npz"), **out)
    print("layers: roi edge", out["roi_edge_out"].shape, "natural", out["roi_natural_out"].shape, "maxpool", out["maxpool_out"].shape)


def main():
    os.makedirs(GOLD, exist_ok=True)
    rtest, rconfig, div, nms = load_reference()
    only = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "--only" else None
    if only in (None, "div"):
        gen_div(div)
    if only in (None, "nms"):
        gen_nms(nms)
    if only in (None, "search"):
        gen_search(rtest, rconfig)
    if only in (None, "blob"):
        gen_blob(rtest, rconfig)
    if only in (None, "tune"):
        gen_tune(rconfig)
    if only in (None, "detect"):
        gen_detect(rtest, rconfig)
    if only in (None, "layers"):
        gen_layers()


if __name__ == "__main__":
    main()
