"""CPU suite: host-side logic of the drop-in layer (config system, shard/gather plumbing with gloo at
world_size 2, blob helpers, capacity/depth planning) -- nothing here needs a GPU."""
import os
import pickle

import numpy as np
import pytest
import torch

from aznet_b200.detect import config as C
from aznet_b200 import dist as azdist
from aznet_b200 import synth


def test_cfg_defaults_and_mode():
    cfg = C.cfg
    assert cfg.TEST.SCALES == (600,) and cfg.TEST.MAX_SIZE in (800, 1000) and cfg.TEST.NMS == 0.5
    assert cfg.SEAR.MIN_SIDE == 10 and cfg.SEAR.NUM_SUBREG == 11 and abs(cfg.DEDUP_BOXES - 1 / 16.) < 1e-12
    assert cfg.EPS == 1e-14 and cfg.SEAR.AZ_CONV == ['conv5_3']
    C.cfg_set_mode('Train')
    assert cfg.SEAR.Tz == 0.0 and cfg.SEAR.NUM_PROPOSALS == 2000
    with pytest.raises(AssertionError):
        C.cfg_set_mode('Test')                      # "testing Tz is not set!"
    C.cfg_set_mode('Test', 0.37)
    assert cfg.SEAR.Tz == 0.37 and cfg.SEAR.NUM_PROPOSALS == 300


def test_cfg_from_file_voc_like(tmp_path):
    yml = tmp_path / "voc.yml"                      # the values of experiments/cfgs/voc.yml
    yml.write_text("TRAIN:\n  IMS_PER_BATCH: 1\n  MAX_SIZE: 800\nTEST:\n  MAX_SIZE: 800\n  NUM_PROPOSALS: 300\n"
                   "SEAR:\n  FIXED_PROPOSAL_NUM: True\n  BATCH_SIZE: 1000\n  AZ_CONV: [conv5_3]\n  FRCNN_CONV: [conv5_3]\n")
    old = (C.cfg.TEST.MAX_SIZE, C.cfg.SEAR.BATCH_SIZE)
    try:
        C.cfg_from_file(str(yml))
        assert C.cfg.TEST.MAX_SIZE == 800 and C.cfg.SEAR.BATCH_SIZE == 1000
        bad = tmp_path / "bad.yml"
        bad.write_text("TEST:\n  NOT_A_KEY: 1\n")
        with pytest.raises(KeyError):
            C.cfg_from_file(str(bad))
        bad.write_text("TEST:\n  MAX_SIZE: big\n")
        with pytest.raises(ValueError):
            C.cfg_from_file(str(bad))
    finally:
        C.cfg.TEST.MAX_SIZE, C.cfg.SEAR.BATCH_SIZE = old


def test_cfg_paths_and_thresh(tmp_path):
    C.cfg_set_path(None)
    assert C.cfg.EXP_DIR == 'default'
    C.cfg_set_path('exp1')

    class I:
        name = 'voc_2007_test'

    class N:
        name = 'vgg16_az'
    assert C.get_output_dir(I, N).endswith(os.path.join('output', 'exp1', 'voc_2007_test', 'vgg16_az'))
    assert C.get_output_dir(I, None).endswith(os.path.join('output', 'exp1', 'voc_2007_test'))
    p = tmp_path / "thresh.pkl"
    pickle.dump(0.42, open(p, "wb"))
    assert C.cfg_load_thresh(str(p)) == 0.42


def test_blob_helpers():
    from aznet_b200.utils.blob import im_list_to_blob, prep_im_for_blob
    a, b = np.ones((4, 6, 3), np.float32), 2 * np.ones((5, 3, 3), np.float32)
    blob = im_list_to_blob([a, b])
    assert blob.shape == (2, 3, 5, 6) and blob.dtype == np.float32
    assert blob[0, :, :4, :6].min() == 1 and blob[0, :, 4:, :].max() == 0 and blob[1, :, :, 3:].max() == 0
    im, s = prep_im_for_blob(np.zeros((600, 1000, 3), np.uint8), np.zeros((1, 1, 3)), 600, 800)
    assert s == 0.8 and im.shape[:2] == (480, 800)


def test_depth_scale_planning():
    from aznet_b200.engine import im_scale_for, search_depth
    assert search_depth(600, 1000, 10) == 6 and search_depth(375, 500, 10) == 6 and search_depth(100, 100, 10) == 4
    assert im_scale_for(600, 1000, (600,), 1000) == 1.0 and im_scale_for(600, 1000, (600,), 800) == 0.8
    assert synth.conv_shape(600, 1000, 1.0) == (38, 63) and synth.conv_shape(600, 1000, 0.8) == (30, 50)


def test_shard_images_partition():
    for n, w in ((1024, 8), (10, 4), (3, 8), (0, 2)):
        parts = [azdist.shard_images(n, r, w) for r in range(w)]
        covered = [i for lo, hi in parts for i in range(lo, hi)]
        assert covered == list(range(n))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_img = 16
    lo, hi = azdist.shard_images(n_img, rank, world)
    P = 5
    boxes = torch.zeros((hi - lo, P, 4), dtype=torch.float64)
    scores = torch.zeros((hi - lo, P), dtype=torch.float32)
    counts = torch.zeros((hi - lo,), dtype=torch.int32)
    for j, img in enumerate(range(lo, hi)):                     # each image's "proposals" encode its index
        counts[j] = 1 + img % P
        boxes[j, :counts[j]] = img
        scores[j, :counts[j]] = img / 100.0
    b, s, c = azdist.gather_proposals(boxes, scores, counts)
    ok = b.shape == (n_img, P, 4) and all(int(c[i]) == 1 + i % P and float(b[i, 0, 0]) == i for i in range(n_img))
    # a run of batches collected per rank and gathered once at the end (bench.py, N > 1)
    col = azdist.ProposalCollector(2, boxes, scores, counts)
    col.add(0, boxes, scores, counts)
    col.add(1, boxes + 1000.0, scores, counts)
    gb, gs, gc = col.gather()
    per = hi - lo
    ok = ok and gb.shape == (world * 2 * per, P, 4)
    for r in range(world):
        first = azdist.shard_images(n_img, r, world)[0]
        ok = ok and float(gb[(r * 2 + 0) * per, 0, 0]) == first and float(gb[(r * 2 + 1) * per, 0, 0]) == first + 1000.0
        ok = ok and int(gc[(r * 2 + 1) * per]) == 1 + first % P
    # several engines (batches in flight), one joint buffer, ONE collective (bench.py --streams)
    grp = azdist.CollectorGroup(2, [(boxes, scores, counts), (boxes, scores, counts)])
    grp.collectors[0].add(0, boxes, scores, counts)
    grp.collectors[1].add(0, boxes + 500.0, scores, counts)
    grp.collectors[1].add(1, boxes + 700.0, scores, counts)
    res = grp.gather()
    ok = ok and len(res) == 2 and res[0][0].shape == (world, 2 * per, P, 4) and res[1][2].shape == (world, 2 * per)
    for r in range(world):
        first = azdist.shard_images(n_img, r, world)[0]
        ok = ok and float(res[0][0][r, 0, 0, 0]) == first and float(res[1][0][r, 0, 0, 0]) == first + 500.0
        ok = ok and float(res[1][0][r, per, 0, 0]) == first + 700.0 and int(res[1][2][r, per]) == 1 + first % P
    # uneven shards: gather_proposals pads every rank's block to the largest shard with count-0 rows
    lo2, hi2 = azdist.shard_images(5, rank, world)
    ub = torch.full((hi2 - lo2, P, 4), float(rank), dtype=torch.float64)
    uc = torch.full((hi2 - lo2,), 2, dtype=torch.int32)
    gb2, _, gc2 = azdist.gather_proposals(ub, torch.zeros((hi2 - lo2, P)), uc)
    per_max = (5 + world - 1) // world
    ok = ok and gb2.shape[0] == world * per_max and int(gc2.sum()) == 2 * 5
    ok = ok and all(int(gc2[r * per_max + k]) == (2 if azdist.shard_images(5, r, world)[0] + k < azdist.shard_images(5, r, world)[1] else 0)
                    for r in range(world) for k in range(per_max))
    # the detection path's exchange step: every rank ends up with the whole set's [imgs, C, mpi] scores / counts
    Cc, mpi = 4, 3
    top = torch.full((hi - lo, Cc, mpi), float("-inf"))
    cnt = torch.zeros((hi - lo, Cc), dtype=torch.int32)
    for j, img in enumerate(range(lo, hi)):
        cnt[j, 1:] = 1 + img % mpi
        top[j, 1:, :int(cnt[j, 1])] = img + 0.5
    t, c2 = azdist.gather_detection_scores(top, cnt)
    ok = ok and t.shape == (n_img, Cc, mpi) and all(float(t[i, 1, 0]) == i + 0.5 and int(c2[i, 2]) == 1 + i % mpi
                                                    for i in range(n_img))
    # the dataset drivers' own sharding (detect/test.py::test_proposals under torch.distributed): contiguous blocks of the
    # imdb's images, per-image lists merged on every rank at the end
    from aznet_b200.detect import test as T
    r, w, mine = T._image_shard(7)
    ok = ok and (r, w) == (rank, world) and list(mine) == (list(range(0, 4)) if rank == 0 else list(range(4, 7)))
    lists = [[] for _ in range(7)]
    for i in mine:
        lists[i] = np.full((1 + i, 4), float(i))
    T._merge_image_lists(lists, mine, w)
    ok = ok and all(isinstance(lists[i], np.ndarray) and lists[i].shape == (1 + i, 4) and float(lists[i][0, 0]) == i for i in range(7))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_proposals_gloo_world2():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(30) for p in procs]
    assert res == [(0, True), (1, True)]


def test_hashnet_is_deterministic():
    net = synth.HashNet(seed=11)
    r = synth.make_rois(50, seed=1)
    z1, p1, d1 = net.heads(r)
    z2, p2, d2 = net.heads(r.copy())
    assert np.array_equal(z1, z2) and np.array_equal(p1, p2) and np.array_equal(d1, d2)
    assert z1.shape == (50, 1) and p1.shape == (50, 11) and d1.shape == (50, 44)
    assert 0 <= p1.min() and p1.max() < 1


def test_host_f32_to_bf16_equals_the_device_rounding():
    """azn_host_f32_to_bf16 (the host-side narrowing of the batched proposal call) == round-to-nearest-even of every
    element, bit for bit, for every thread count and for lengths / alignments that exercise the vector body and both
    scalar tails; NaN -> 0x7fff like __float2bfloat16_rn.  A host function: no GPU involved."""
    from aznet_b200 import _lib, ops
    _lib.build()
    assert _lib.lib().azn_host_threads() >= 1
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal(300007).astype(np.float32))
    bits = torch.from_numpy(rng.integers(0, 2 ** 32, 4096, dtype=np.uint64).astype(np.uint32).view(np.float32).copy())   # any bit pattern
    special = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 1e-40, -1e-40, 3.4e38, -3.4e38, 1.00390625, 1.01171875,
                            float("nan")], dtype=torch.float32)
    x = torch.cat([special, bits, x])
    want = x.to(torch.bfloat16).view(torch.int16)
    nan = torch.isnan(x)
    for threads in (0, 1, 3, 8):
        for off, n in ((0, x.numel()), (1, 4099), (3, 17), (5, 15), (7, 1), (2, 0), (11, 65536 + 5)):
            src = x[off:off + n].clone()
            buf = torch.zeros(n + 16, dtype=torch.bfloat16)
            dst = buf[off % 16:off % 16 + n]                       # outputs at every alignment
            ops.host_f32_to_bf16(src, dst, threads)
            got = dst.view(torch.int16)
            m = ~nan[off:off + n]
            assert torch.equal(got[m], want[off:off + n][m]), (threads, off, n)
            assert bool(((got[~m] & 0x7fff) == 0x7fff).all()), "NaN pattern"
            assert int(buf[off % 16 + n:].view(torch.int16).abs().sum()) == 0 and int(buf[:off % 16].view(torch.int16).abs().sum()) == 0, "wrote outside"
    with pytest.raises(AssertionError):
        ops.host_f32_to_bf16(x, torch.zeros(3, dtype=torch.bfloat16))


def test_reference_arm_line_contract():
    """`bench.py --impl reference` (the arm the driver runs beside the GPU arm) needs no GPU: one JSON line with the GPU arm's
    metric / unit / config object (bench.workload_config, key for key), `impl`, its own cpu_baseline and an e2e object
    without copies."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--no-extra"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-800:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    sys.path.insert(0, root)
    import bench
    assert line["impl"] == "reference" and line["metric"] == "AZ proposal images/sec" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["config"] == bench.workload_config(1)
    assert set(bench.workload_config(2, 1024)) == set(bench.workload_config(1)) | {"job"}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
