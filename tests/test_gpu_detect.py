"""GPU suite: the batched detection step (azn_detect_* + azn_nms_segments) against the oracle's restatement of
_frcnn_forward + test_net + apply_nms, driven by the integer-hash detection net so that every selection is
pinned bit for bit; and the real tensor-core head against the fp32 oracle within the stated tolerance."""
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


class _FakeHead:
    """Dimensions only: DetectEngine with externally supplied head outputs."""
    pooled, C = 7, 8

    def __init__(self, dev, num_classes):
        self.num_classes = num_classes
        self.w6 = torch.zeros((64, 7 * 7 * 8), dtype=torch.bfloat16, device=dev)
        self.w7 = torch.zeros((64, 64), dtype=torch.bfloat16, device=dev)


def _proposals(n_img, cap, H, W, counts, seed):
    boxes = np.zeros((n_img, cap, 4))
    for i, c in enumerate(counts):
        b = synth.make_boxes(c, H, W, seed=seed + i, lo=12, hi=500)
        if c > 40:
            b[30:40] = b[5:15] + 0.25                   # same feature-space cell, different image-space box
            b[c - 3:] = b[:3]                            # exact duplicates
        boxes[i, :c] = b
    return boxes


@pytest.mark.parametrize("num_classes,H,W,max_size,bs,counts", [
    (21, 600, 1000, 1000, 10000, [300, 257, 300]),      # VOC-like: max_per_set = 40 * 3 < pushed -> thresholds active
    (81, 480, 640, 800, 10000, [300, 300, 12, 0]),       # COCO-like: max_per_set = 10 * 4; an image without proposals
    (21, 375, 500, 1000, 100, [300, 150]),               # chunked: dedup per BATCH_SIZE = 100 boxes
])
def test_detection_step_matches_oracle(dev, O, num_classes, H, W, max_size, bs, counts):
    from aznet_b200 import detector
    n_img, cap = len(counts), 300
    boxes = _proposals(n_img, cap, H, W, counts, seed=21)
    eng = detector.DetectEngine(_FakeHead(dev, num_classes), n_img, H, W, cap, max_size=max_size, batch_size=bs)
    net = synth.HashDetNet(seed=13, num_classes=num_classes)
    boxes_d = torch.from_numpy(boxes).to(dev)
    counts_d = torch.tensor(counts, dtype=torch.int32, device=dev)
    eng.prepare(boxes_d, counts_d)
    m = int(eng.m_total.item())
    rois = eng.rois[:m].cpu().numpy().copy()
    rois[:, 0] = 0                                        # the reference's per-image blob has level column 0
    p, d = net.heads(rois)
    head = np.zeros((n_img * cap, eng.ld), np.float32)
    head[:m, :num_classes], head[:m, num_classes:5 * num_classes] = p, d
    dets, tops, cnt = eng.select(head_out=torch.from_numpy(head).to(dev))
    # oracle
    cfg = O.OracleCfg(TEST_MAX_SIZE=max_size, BATCH_SIZE=bs)
    per_image, n_uniq = [], []
    for i, c in enumerate(counts):
        if c == 0:
            per_image.append(None)
            continue
        s_ref, p_ref, _ = O.frcnn_forward({"full": net, "fc": net}, (H, W, 3), boxes[i, :c], num_classes,
                                          {"conv5_3": np.zeros((1, 1, 2, 2), np.float32)}, cfg)
        per_image.append((s_ref, p_ref))
    ref_boxes, ref_thresh = O.test_net_select(per_image, num_classes)
    max_per_set = 800 // (num_classes - 1) * n_img
    assert np.isfinite(ref_thresh[1:]).any(), "the case must exercise the threshold"
    pre = detector.detections_to_host(dets, cnt)          # before the final filter: superset of the reference's rows
    thresh, keep, keep_count = detector.finish_detections(dets, tops, cnt, max_per_set, 0.3)
    np.testing.assert_array_equal(thresh.cpu().numpy()[1:], ref_thresh[1:].astype(np.float32))
    got = detector.detections_to_host(dets, cnt)
    for j in range(1, num_classes):
        for i in range(n_img):
            r = ref_boxes[j][i]
            if per_image[i] is None:
                assert len(got[j][i]) == 0
                continue
            g = got[j][i]
            assert g.shape == r.shape, (j, i, g.shape, r.shape)
            np.testing.assert_array_equal(g[:, 4], r[:, 4])                    # scores: bit-exact, same order
            np.testing.assert_allclose(g[:, :4], r[:, :4], rtol=1e-5, atol=1e-4)  # float32 exp ulps only
            assert len(pre[j][i]) >= len(g)
    # NMS: the kept rows of every (class, image) equal the oracle's apply_nms on the same float32 detections
    ref_nms = O.apply_nms(got, 0.3)
    post = detector.detections_to_host(dets, cnt, keep, keep_count)
    for j in range(1, num_classes):
        for i in range(n_img):
            a, b = post[j][i], ref_nms[j][i]
            assert (len(a) == 0 and len(b) == 0) or np.array_equal(a, b), (j, i)


def test_detect_thresholds_streamed_equals_one_shot(dev):
    """Order independence of the heap cap: thresholds from two half sets concatenated == from the whole set."""
    from aznet_b200 import ops
    rng = np.random.default_rng(5)
    N, C, mpi = 16, 7, 100
    cnt = rng.integers(0, mpi + 1, (N, C)).astype(np.int32)
    top = -np.sort(-rng.random((N, C, mpi), dtype=np.float32), axis=2)
    t = ops.detect_thresholds(torch.from_numpy(top).to(dev), torch.from_numpy(cnt).to(dev), 300).cpu().numpy()
    for j in range(1, C):
        vals = np.concatenate([top[i, j, :cnt[i, j]] for i in range(N)])
        want = -np.inf if len(vals) <= 300 else np.sort(vals)[::-1][299]
        assert t[j] == np.float32(want), (j, t[j], want)
    assert t[0] == -np.inf


def test_detect_engine_real_head_vs_oracle(dev, O):
    """bf16 tensor-core Fast R-CNN head vs the fp32 oracle on bf16-rounded weights: class probabilities within
    2e-2, and the selected detections of the confident classes agree."""
    from aznet_b200 import detector, engine, ops
    from aznet_b200.net import FRCNNHeadWeights
    Cc, H, W, n_img, ncls, cap = 64, 240, 320, 2, 6, 128
    w = synth.make_frcnn_weights(seed=4, num_classes=ncls, C=Cc, h6=256, h7=256)
    s = engine.im_scale_for(H, W)
    fh, fw = synth.conv_shape(H, W, s)
    conv = synth.make_conv_maps(n_img, Cc, fh, fw, seed=9)
    bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    wq = {k: (bf(v[0]), v[1]) for k, v in w.items()}
    head = FRCNNHeadWeights(w, dev)
    eng = detector.DetectEngine(head, n_img, H, W, cap)
    counts = [128, 77]
    boxes = _proposals(n_img, cap, H, W, counts, seed=33)
    nhwc = ops.nchw_to_nhwc_bf16(torch.from_numpy(conv).to(dev))
    dets, tops, cnt = eng.detect(nhwc, torch.from_numpy(boxes).to(dev), torch.tensor(counts, dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    out = eng.out.cpu().numpy()
    inv, off = eng.inv.cpu().numpy(), eng.img_off.cpu().numpy()
    cfg = O.OracleCfg()
    for i, c in enumerate(counts):
        net = O.OracleNet(wq, "frcnn", cfg=cfg)
        s_ref, p_ref, _ = O.frcnn_forward({"full": net, "fc": net}, (H, W, 3), boxes[i, :c], ncls, {"conv5_3": bf(conv[i:i + 1])}, cfg)
        rows = off[i] + inv[i, :c]
        np.testing.assert_allclose(out[rows, :ncls], s_ref, atol=2e-2)
        np.testing.assert_allclose(out[rows, :ncls].sum(1), 1.0, atol=1e-4)
        d = dets[i].cpu().numpy()
        n = cnt[i].cpu().numpy()
        for j in range(1, ncls):
            assert n[j] == min(c, 100)
            top_ref = np.sort(s_ref[:, j])[::-1][:n[j]]
            np.testing.assert_allclose(d[j, :n[j], 4], top_ref, atol=2e-2)
            assert (np.diff(d[j, :n[j], 4]) <= 0).all()
