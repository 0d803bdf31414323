"""GPU suite: the fused level kernel + selection against (a) the golden vectors made by the
reference's own im_propose with the integer-hash net, (b) the oracle with the real fc heads."""
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
    """Head dimensions only (no weights): lets SearchEngine run with externally supplied head outputs."""
    nsub, pooled, C, h6, h71, h72 = 11, 7, 8, 64, 64, 64
    n_head, ld_head = 56, 56

    def __init__(self, dev):
        self.w6 = torch.zeros((64, 7 * 7 * 8), dtype=torch.bfloat16, device=dev)


def _host_heads(eng, net):
    m = int(eng.m_total.item())
    rois = eng.rois[:m].cpu().numpy().copy()
    rois[:, 0] = 0                                      # HashNet hashes the reference's per-image blob (level column 0)
    z, p, d = net.heads(rois)
    h = np.zeros((m, 56), np.float32)
    h[:, :11], h[:, 11:55], h[:, 55] = p, d, z[:, 0]
    eng.heads[:m] = torch.from_numpy(h).to(eng.heads.device)


def _drive_with_hashnet(eng, net, n_img):
    """Level loop with the heads computed on the host by HashNet from the ROIs the GPU produced; the same
    launch sequence as SearchEngine.propose (levels 1+2 in one pass of the heads when eng.merge_root)."""
    first = 1
    if eng.merge_root:
        eng.begin_merged()
        _host_heads(eng, net)
        eng.search_level(1, root_props=True)
        eng.search_level(2)
        first = 3
    else:
        eng.begin()
    for k in range(first, eng.n_levels + 1):
        _host_heads(eng, net)
        eng.search_level(k)
    eng.select()
    return eng.results()


CASES = ["d0_600x1000", "voc_600x1000", "fullzoom_600x1000", "small_375x500", "chunked_480x640",
         "tc_thresh_333x500", "nozoom_600x1000"]


@pytest.mark.parametrize("merge_root", [True, False], ids=["merged12", "level_by_level"])
@pytest.mark.parametrize("name", CASES)
def test_search_matches_reference_golden(dev, golden, name, merge_root):
    from aznet_b200 import engine
    g = golden["search"]
    H, W, max_size, bs, tz, rate, nprop, fixed = g[name + "_cfg"]
    n_img = 3
    eng = engine.SearchEngine(_FakeHead(dev), n_img, int(H), int(W), max_size=int(max_size), batch_size=int(bs), tz=float(tz),
                              fixed_num=bool(fixed), num_proposals=300 if nprop < 0 else int(nprop), merge_root=merge_root)
    boxes, scores, n_eval, depth = _drive_with_hashnet(eng, synth.HashNet(seed=11, zoom_rate=float(rate)), n_img)
    ref = g[name + "_Y"]
    log = str(g[name + "_log"])
    for i in range(n_img):                              # every image of the batch is the same problem
        assert "{0} proposals, evaluate {1} regions, reaches depth {2}.".format(len(boxes[i]), n_eval[i], depth[i]) == log
        if fixed:
            # same rows in the same order unless scores tie; decode differs only by float32 exp ulps
            order_ok = np.allclose(boxes[i], ref, rtol=1e-5, atol=1e-5)
            if not order_ok:
                a = boxes[i][np.lexsort(boxes[i].T[::-1])]
                b = ref[np.lexsort(ref.T[::-1])]
                np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5)
        else:
            np.testing.assert_allclose(boxes[i], ref, rtol=1e-5, atol=1e-5)


TUNE_CASES = ["train_375x500", "train_voc_600x1000", "tz05_480x640", "stop_early_333x500"]


@pytest.mark.parametrize("name", TUNE_CASES)
def test_tune_search_matches_reference_golden(dev, golden, name):
    """SearchEngine(tune=True) = the reference's diagnostic im_propose (lib/detect/tune.py:256-316): K levels, Tz = 0
    at the first, anchor history bit for bit (regions f64, zoom scores f32-exact)."""
    from aznet_b200 import engine
    g = golden["tune"]
    H, W, max_size, bs, tz, rate, nprop = g[name + "_cfg"]
    n_img = 2
    eng = engine.SearchEngine(_FakeHead(dev), n_img, int(H), int(W), max_size=int(max_size), batch_size=int(bs), tz=float(tz),
                              fixed_num=True, num_proposals=int(nprop), tune=True)
    assert eng.n_levels == eng.K and not eng.merge_root
    boxes, scores, n_eval, depth = _drive_with_hashnet(eng, synth.HashNet(seed=11, zoom_rate=float(rate)), n_img)
    hist = eng.history()
    ref_h, ref_y = g[name + "_Bhis"], g[name + "_Y5"]
    for i in range(n_img):
        assert "{0} proposals, evaluate {1} regions, reaches depth {2}.".format(len(boxes[i]), n_eval[i], depth[i] - 1) == str(g[name + "_log"])
        assert np.array_equal(hist[i].view(np.uint64), ref_h.view(np.uint64)), "anchor history differs"
        y5 = np.hstack((boxes[i], scores[i][:, None].astype(np.float64)))
        if not np.allclose(y5, ref_y, rtol=1e-5, atol=1e-5):
            np.testing.assert_allclose(y5[np.lexsort(y5.T[::-1])], ref_y[np.lexsort(ref_y.T[::-1])], rtol=1e-5, atol=1e-5)


def test_tune_threshold_kernel(dev, golden):
    """azn_tune_threshold = the reference's heap: k-th highest zoom score, -inf when fewer anchors, ragged counts."""
    from aznet_b200 import ops
    rng = np.random.default_rng(5)
    n, cap = 37, 3000
    counts = rng.integers(0, cap + 1, n).astype(np.int32)
    counts[3] = 0
    z = rng.random((n, cap), dtype=np.float32)
    z[5, :10] = z[6, :10]                                                # ties across images
    flat = np.sort(np.concatenate([z[i, :counts[i]] for i in range(n)]))[::-1]
    zt, ct = torch.from_numpy(z).to(dev), torch.from_numpy(counts).to(dev)
    for k in (1, 20 * n, len(flat) - 1):
        assert float(ops.tune_threshold(zt, ct, k).item()) == float(flat[k - 1])
    assert float(ops.tune_threshold(zt, ct, len(flat)).item()) == -np.inf
    assert float(ops.tune_threshold(zt, ct, len(flat) + 5).item()) == -np.inf


def test_search_levels_match_oracle_trace(dev, O):
    """Regions of every level (order included) are bit-identical to the oracle's B."""
    from aznet_b200 import engine
    H, W = 600, 1000
    net = synth.HashNet(seed=11, zoom_rate=0.5)
    cfg = O.OracleCfg(Tz=0.5)
    trace = []
    O.im_propose({"full": net, "fc": net}, (H, W, 3), cfg, conv={"conv5_3": np.zeros((1, 1, 2, 2), np.float32)}, trace=trace)
    eng = engine.SearchEngine(_FakeHead(dev), 1, H, W, tz=0.5, merge_root=False)
    eng.begin()
    for k in range(1, eng.n_levels + 1):
        n = int(eng.n_regions[eng._cur][0].item())
        B = eng.regions[eng._cur][0, :n].cpu().numpy()
        assert np.array_equal(B.view(np.uint64), trace[k - 1]["B"].view(np.uint64)), "level %d regions differ" % k
        m = int(eng.m_total.item())
        rois = eng.rois[:m].cpu().numpy().copy()
        rois[:, 0] = 0
        z, p, d = net.heads(rois)
        h = np.zeros((m, 56), np.float32)
        h[:, :11], h[:, 11:55], h[:, 55] = p, d, z[:, 0]
        eng.heads[:m] = torch.from_numpy(h).to(dev)
        eng.search_level(k)
    torch.cuda.synchronize()
    assert int(eng.status.item()) == 0


def test_search_capacity_overflow_is_reported(dev):
    from aznet_b200 import engine
    eng = engine.SearchEngine(_FakeHead(dev), 1, 600, 1000, tz=0.0)
    eng._st.cap_props = 5                                # lie about the capacity: the kernel must flag it, not overrun
    with pytest.raises(RuntimeError, match="capacity"):
        _drive_with_hashnet(eng, synth.HashNet(seed=11), 1)


def test_engine_real_heads_vs_oracle(dev, O):
    """bf16 tensor-core heads vs the fp32 oracle on the same (bf16-rounded) weights and maps: head outputs
    within 2e-2 on level 1, and proposal-set recall parity for the whole search."""
    from aznet_b200 import engine, ops
    C, H, W, n_img = 64, 375, 500, 4
    w = synth.make_az_weights(seed=3, C=C, h6=512, h71=192, h72=64, zoom_bias=-0.3)
    s = engine.im_scale_for(H, W)
    fh, fw = synth.conv_shape(H, W, s)
    conv = synth.make_conv_maps(n_img, C, fh, fw, seed=7)
    bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    wq = {k: (bf(v[0]), v[1]) for k, v in w.items()}
    head = engine.AZHeadWeights(w, dev)
    eng = engine.SearchEngine(head, n_img, H, W, num_proposals=300, tz=0.5)
    nhwc = ops.nchw_to_nhwc_bf16(torch.from_numpy(conv).to(dev))
    # level-1 head outputs (the root rows of the merged pass are the same rows 0 .. n_img-1)
    eng.begin()
    eng.run_heads(nhwc, 1)
    torch.cuda.synchronize()
    heads = eng.heads[:n_img].cpu().numpy()
    cfg = O.OracleCfg(NUM_PROPOSALS=300, Tz=0.5)
    for i in range(n_img):
        net = O.OracleNet(wq, "az", cfg=cfg)
        roi = np.array([[0, 0, 0, (W - 1) * s, (H - 1) * s]], np.float32)
        net.blobs["rois"].reshape(1, 5)
        net.blobs["conv5_3"].reshape(1, C, fh, fw)
        out = net.forward(rois=roi, conv5_3=bf(conv[i:i + 1]))
        np.testing.assert_allclose(heads[i, :11], out["adj_prob"][0], atol=2e-2)
        np.testing.assert_allclose(heads[i, 11:55], out["adj_bbox"][0], atol=2e-2)
        np.testing.assert_allclose(heads[i, 55], out["zoom_prob"][0, 0], atol=2e-2)
    # whole search
    eng.propose(nhwc)
    boxes, scores, n_eval, depth = eng.results()

    def iou(a, b):
        x1, y1 = np.maximum(a[:, None, 0], b[None, :, 0]), np.maximum(a[:, None, 1], b[None, :, 1])
        x2, y2 = np.minimum(a[:, None, 2], b[None, :, 2]), np.minimum(a[:, None, 3], b[None, :, 3])
        inter = np.clip(x2 - x1 + 1, 0, None) * np.clip(y2 - y1 + 1, 0, None)
        aa = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
        ab = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        return inter / (aa[:, None] + ab[None, :] - inter)

    for i in range(n_img):
        net = O.OracleNet(wq, "az", cfg=cfg)
        Y, sc, info = O.im_propose({"full": net, "fc": net}, (H, W, 3), cfg, conv={"conv5_3": bf(conv[i:i + 1])},
                                   return_scores=True)
        assert abs(int(n_eval[i]) - info["num_eval"]) <= max(3, 0.05 * info["num_eval"])
        # recall parity: the oracle's proposals are recovered by ours at IoU >= 0.9 (and vice versa)
        m = iou(Y, boxes[i])
        assert (m.max(1) >= 0.9).mean() >= 0.95, (m.max(1) >= 0.9).mean()
        assert (m.max(0) >= 0.9).mean() >= 0.95


def test_collector_slots_follow_the_device_counter(dev):
    """azn_collect_proposals: the final lists of batch k land in slot k % n_slots of the per-rank collection, the
    slot index comes from the device-side counter (same launch every step, also under CUDA-graph replay)."""
    from aznet_b200 import engine, ops
    from aznet_b200.dist import ProposalCollector
    C, H, W, n_img = 64, 200, 320, 3
    w = synth.make_az_weights(seed=3, C=C, h6=256, h71=64, h72=64, zoom_bias=0.2)
    head = engine.AZHeadWeights(w, dev)
    eng = engine.SearchEngine(head, n_img, H, W, num_proposals=50, tz=0.5)
    fh, fw = synth.conv_shape(H, W, eng.scale)
    maps = [ops.nchw_to_nhwc_bf16(torch.from_numpy(synth.make_conv_maps(n_img, C, fh, fw, seed=7 + k)).to(dev)) for k in range(2)]
    want = []
    for k in range(2):
        eng.propose(maps[k])
        torch.cuda.synchronize()
        want.append((eng.out_boxes.clone(), eng.out_scores.clone(), eng.out_count.clone()))
    col = ProposalCollector(3, eng.out_boxes, eng.out_scores, eng.out_count)
    eng.collector = col
    graphs = [eng.capture(m)[0] for m in maps]          # capture runs warm-up passes: the counter has moved
    col.reset()
    for k in (0, 1, 1, 0):                               # batches 0..3 -> slots 0, 1, 2, 0
        graphs[k].replay()
    torch.cuda.synchronize()
    assert int(col.state[0].item()) == 4 and int(col.state[1].item()) == 0
    for slot, k in ((0, 0), (1, 1), (2, 1)):
        assert torch.equal(col.boxes[slot], want[k][0]) and torch.equal(col.scores[slot], want[k][1])
        assert torch.equal(col.counts[slot], want[k][2])
    gb, gs, gc = col.gather()                            # single process: the flattened collection itself
    assert gb.shape[0] == 3 * n_img and torch.equal(gc[:n_img], want[0][2])


@pytest.mark.parametrize("host_narrow", ["off", "on", "auto", "split"])
def test_proposal_pipeline_routes_agree_bit_for_bit(dev, host_narrow):
    """The batched host-facing call (f32 NCHW host maps in, proposal lists out): uploading the f32 batch, or rounding it
    to bf16 on the host first (azn_host_f32_to_bf16 + azn_nchw_bf16_to_nhwc_bf16), or splitting the batch between the two
    (some images raw, the others narrowed meanwhile), puts the same bits in HBM, so every
    route returns exactly what the engine returns on the resident bf16 map -- with and without CUDA-graph replay,
    and again when the slots are reused."""
    from aznet_b200 import engine, ops
    from aznet_b200.pipeline import ProposalPipeline
    C, H, W, n_img = 64, 375, 500, 4
    w = synth.make_az_weights(seed=3, C=C, h6=512, h71=192, h72=64, zoom_bias=-0.3)
    s = engine.im_scale_for(H, W)
    fh, fw = synth.conv_shape(H, W, s)
    head = engine.AZHeadWeights(w, dev)
    eng = engine.SearchEngine(head, n_img, H, W, num_proposals=300, tz=0.5)
    batches = [synth.make_conv_maps(n_img, C, fh, fw, seed=7 + k) for k in range(3)]
    batches[1][0, 0, 0, :4] = [np.nan, -0.0, 65504.0, 1e-40]         # NaN never wins a max; no inf (inf - inf in the heads)
    want = []
    for conv in batches:
        dev_f32 = torch.from_numpy(conv).to(dev)
        nhwc = ops.nchw_to_nhwc_bf16(dev_f32)
        via16 = ops.nchw_bf16_to_nhwc_bf16(dev_f32.to(torch.bfloat16))
        assert torch.equal(nhwc.view(torch.int16), via16.view(torch.int16))
        eng.propose(nhwc)
        torch.cuda.synchronize()
        want.append([t.clone() for t in (eng.out_boxes, eng.out_scores, eng.out_count, eng.n_eval)])
    def same(got, ref, tag):
        boxes, scores, count, n_eval = [t.cpu() for t in ref]
        assert torch.equal(got[2], count) and torch.equal(got[3], n_eval), tag
        for i in range(n_img):                                      # rows past an image's count are not part of the result
            c = int(count[i])
            assert torch.equal(got[0][i, :c].view(torch.int64), boxes[i, :c].view(torch.int64)), tag          # bit patterns
            assert torch.equal(got[1][i, :c].view(torch.int32), scores[i, :c].view(torch.int32)), tag

    for use_graph in (False, True):
        pipe = ProposalPipeline(eng, batches[0].shape, depth=2, use_graph=use_graph, host_narrow=host_narrow, host_threads=3, narrow_chunks=3)
        hosts = [torch.from_numpy(b).pin_memory() for b in batches]
        tickets = []
        for k in (0, 1, 2, 1, 0):
            tickets.append((k, pipe.submit(hosts[k])))
            if len(tickets) == 2:
                kk, t = tickets.pop(0)
                same(pipe.result(t), want[kk], (host_narrow, use_graph, kk))
        kk, t = tickets.pop(0)
        same(pipe.result(t), want[kk], (host_narrow, use_graph, kk))
        assert pipe.narrow == (host_narrow in ("on", "split")) or host_narrow == "auto"
        per_img = batches[0].size // n_img
        raw = pipe.raw_images if pipe.narrow else n_img               # images that cross the link as f32
        assert pipe.h2d_bytes == per_img * (4 * raw + 2 * (n_img - raw))
        if host_narrow == "split":
            assert 0 < pipe.raw_images < n_img
        if host_narrow == "auto":
            assert pipe.narrow_timing and (pipe.narrow_timing["chosen"] in ("f32_upload", "host_bf16_then_upload")
                                           or pipe.narrow_timing["chosen"].startswith("split_"))
