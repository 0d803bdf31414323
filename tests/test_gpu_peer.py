"""Peer window (the gather of the per-image proposal lists): the ranks' azn_collect_proposals append
straight into rank 0's memory.  Two processes on ONE GPU (CUDA IPC works between processes on the same device;
gloo carries the handle and the barrier), so the test runs on the single-GPU box; bench.py --gpus N exercises the
same code over NVLink with NCCL."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_workers(procs, q, timeout_s):
    """Start the workers and collect one result each; the first failure (or a time-out) ends the others at once -- a
    rank that died must not leave its peer waiting in a collective for the rest of the GPU lease."""
    import queue
    import time
    [p.start() for p in procs]
    res, deadline = [], time.monotonic() + timeout_s
    try:
        while len(res) < len(procs) and time.monotonic() < deadline:
            try:
                r = q.get(timeout=2.0)
            except queue.Empty:
                if all(not p.is_alive() for p in procs) and q.empty():
                    break
                continue
            res.append(r)
            if not r[1]:
                break
    finally:
        for p in procs:
            p.join(5 if len(res) == len(procs) and all(r[1] for r in res) else 0.1)
            if p.is_alive():
                p.terminate()
                p.join(5)
    return sorted(res)


def _peer_worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from aznet_b200 import _lib, dist as azdist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        _lib.require_device()
        n_img, P, slots = 4, 6, 3
        mk = lambda k: (torch.full((n_img, P, 4), 1000.0 * rank + k, dtype=torch.float64, device=dev),
                        torch.full((n_img, P), 0.5 * rank + k, dtype=torch.float32, device=dev),
                        torch.full((n_img,), 1 + rank + k, dtype=torch.int32, device=dev))
        b0, s0, c0 = mk(0)
        grp = azdist.CollectorGroup(slots, [(b0, s0, c0), (b0, s0, c0)], peer=True)
        ok = grp.window is not None
        grp.reset()
        # engine 0 appends batches 0..3 (slot 0 is overwritten by batch 3), engine 1 one batch
        for k in range(4):
            grp.collectors[0].device_add(*mk(k))
        grp.collectors[1].device_add(*mk(7))
        res = grp.gather()
        torch.cuda.synchronize()
        if rank == 0:
            for r in range(world):
                b, s, c = res[0]
                ok = ok and b.shape == (world, slots * n_img, P, 4) and c.shape == (world, slots * n_img)
                for slot, k in ((0, 3), (1, 1), (2, 2)):
                    ok = ok and float(b[r, slot * n_img, 0, 0]) == 1000.0 * r + k and float(b[r, slot * n_img + n_img - 1, P - 1, 3]) == 1000.0 * r + k
                    ok = ok and float(s[r, slot * n_img + 1, 2]) == 0.5 * r + k and int(c[r, slot * n_img + 2]) == 1 + r + k
                b1, _, c1 = res[1]
                ok = ok and float(b1[r, 0, 0, 0]) == 1000.0 * r + 7 and int(c1[r, n_img - 1]) == 8 + r and int(c1[r, n_img]) == 0
        try:
            grp.collectors[0].add(0, b0, s0, c0)
            ok = False
        except RuntimeError:
            pass
        dist.barrier()
        grp.window.close()
        q.put((rank, bool(ok), ""))
        dist.destroy_process_group()
    except Exception as e:                                   # the parent must not wait for a dead child
        q.put((rank, False, repr(e)))


def test_peer_window_append_two_processes_one_gpu():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    res = _run_workers(procs, q, 120)
    assert res == [(0, True, ""), (1, True, "")], res


def _small_az_nets(dev, num_classes=6):
    import numpy as np
    from aznet_b200 import backbone, net, synth
    bw = backbone.make_vgg16_weights(seed=5, width_div=8)                 # conv5_3 has 64 channels
    bb = backbone.VGG16Backbone(bw, dev)
    azw = synth.make_az_weights(seed=3, C=64, h6=256, h71=96, h72=32, zoom_bias=0.0)
    frw = synth.make_frcnn_weights(seed=4, num_classes=num_classes, C=64, h6=256, h7=128)
    frw["cls_score"] = (frw["cls_score"][0] * np.float32(0.02), frw["cls_score"][1])      # O(1) logits: unsaturated softmax
    frw["bbox_pred"] = (frw["bbox_pred"][0] * np.float32(0.05), frw["bbox_pred"][1])
    az = {"full": net.Net(azw, "az", backbone=bb, name="az_small"), "fc": net.Net(azw, "az", name="az_small")}
    fr = {"full": net.Net(frw, "frcnn", backbone=bb, name="frcnn_small"), "fc": net.Net(frw, "frcnn", name="frcnn_small")}
    return az, fr


def _shard_worker(rank, world, port, root_dir, q):
    """test_proposals under torch.distributed: every rank runs the driver on the same imdb; the images are sharded,
    the lists merged, rank 0 writes proposals.pkl (the reference's one file)."""
    try:
        import contextlib
        import io
        import pickle
        import numpy as np
        import torch.distributed as dist
        from aznet_b200 import _lib, synth
        from aznet_b200.detect import config as C
        from aznet_b200.detect import test as T
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        _lib.require_device()
        C.cfg_set_path("pytest_shard")
        C.cfg_set_mode("Test", 0.5)
        C.cfg.ROOT_DIR = root_dir
        az, fr = _small_az_nets(dev)
        ims = synth.make_images(5, 200, 300, seed=80) + synth.make_images(2, 240, 200, seed=90)
        ims = [ims[k] for k in (0, 5, 1, 2, 6, 3, 4)]
        buf = io.StringIO()
        single, diag = None, ""

        def close(multi_d, single_d):
            """all_boxes[cls][img] of two runs: nearly every detection of one has a partner in the other (different batch
            compositions flip a few bf16 roundings: boxes to 0.5 px, scores to 0.02)."""
            hit = tot = 0
            for j in range(1, 6):
                for i in range(7):
                    a, b = np.asarray(multi_d[j][i]).reshape(-1, 5), np.asarray(single_d[j][i]).reshape(-1, 5)
                    tot += len(b)
                    hit += sum(1 for row in b if len(a) and (np.abs(a[:, :4] - row[:4]).max(axis=1) + 25.0 * np.abs(a[:, 4] - row[4])).min() <= 0.5)
            return tot, hit

        if rank == 0:                                            # the single-process results first (no process group yet)
            mem = synth.InMemoryImdb(ims, num_classes=6, name="shard_single")
            with contextlib.redirect_stdout(buf):
                T.test_proposals(az, mem)
                single_file = os.path.join(C.get_output_dir(mem, az["full"]), "proposals.pkl")
                T.test_net(fr, single_file, mem)
                single_det = pickle.load(open(os.path.join(C.get_output_dir(mem, fr["full"]), "detections.pkl"), "rb"))
                T.test_net_shared(az, fr, mem)
                single_shared = pickle.load(open(os.path.join(C.get_output_dir(mem, az["full"]), "detections.pkl"), "rb"))
            single = pickle.load(open(single_file, "rb"))["boxes"]
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        import datetime
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=90))
        imdb = synth.InMemoryImdb(ims, num_classes=6, name="shard_multi")
        with contextlib.redirect_stdout(buf):
            T.test_proposals(az, imdb)
        st = T.test_proposals.last_stats
        ok = st["world"] == world and st["rank"] == rank and st["images"] == (4 if rank == 0 else 3) and st["route"] == "batched"
        dist.barrier()
        path = os.path.join(C.get_output_dir(imdb, az["full"]), "proposals.pkl")
        if rank == 0:
            multi = pickle.load(open(path, "rb"))["boxes"]
            ok = ok and len(multi) == 7
            hit = tot = 0
            for i in range(7):
                a, b = np.asarray(multi[i]), np.asarray(single[i])
                ok = ok and a.dtype == np.float64 and 0 < a.shape[0] <= 300
                # different batch compositions (4 + 3 images instead of 5 + 2): bf16 flips may move a few boxes
                tot += len(b)
                hit += sum(1 for row in b if len(a) and np.abs(a - row).max(axis=1).min() <= 0.05)
            diag = "" if (ok and tot >= 150 and hit >= 0.97 * tot) else "stats %r ok %r rows %d matched %d counts %r" % (
                dict(st), ok, tot, hit, [(len(multi[i]), len(single[i])) for i in range(7)])
            ok = ok and tot >= 150 and hit >= 0.97 * tot
        dist.barrier()
        # detection over the saved proposals, then the shared-conv route: sharded images, set-wide thresholds (one exchange
        # of the scores), merged nesting, one detections.pkl written by rank 0
        with contextlib.redirect_stdout(buf):
            T.test_net(fr, path, imdb)
        st2 = dict(T.test_net.last_stats)
        dist.barrier()
        if rank == 0:
            multi_det = pickle.load(open(os.path.join(C.get_output_dir(imdb, fr["full"]), "detections.pkl"), "rb"))
            tot, hit = close(multi_det, single_det)
            good = st2.get("world") == world and st2["route"] == "batched" and len(multi_det) == 6 and len(multi_det[1]) == 7 and tot >= 100 and hit >= 0.93 * tot
            if not good:
                diag += " | test_net: stats %r rows %d matched %d" % (st2, tot, hit)
            ok = ok and good
        dist.barrier()
        with contextlib.redirect_stdout(buf):
            T.test_net_shared(az, fr, imdb)
        st3 = dict(T.test_net_shared.last_stats)
        dist.barrier()
        if rank == 0:
            multi_shared = pickle.load(open(os.path.join(C.get_output_dir(imdb, az["full"]), "detections.pkl"), "rb"))
            tot, hit = close(multi_shared, single_shared)
            good = st3.get("world") == world and st3["route"] == "batched" and tot >= 100 and hit >= 0.93 * tot
            if not good:
                diag += " | test_net_shared: stats %r rows %d matched %d" % (st3, tot, hit)
            ok = ok and good
        dist.barrier()
        q.put((rank, bool(ok), diag if rank == 0 else ""))
        dist.destroy_process_group()
    except Exception as e:
        import traceback
        q.put((rank, False, traceback.format_exc()[-600:]))


def test_dataset_drivers_shard_images_over_ranks(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    res = _run_workers(procs, q, 200)
    assert res == [(0, True, ""), (1, True, "")], res
