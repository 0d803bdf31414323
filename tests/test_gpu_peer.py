"""Peer window (the gather of the per-image proposal lists): the ranks' azn_collect_proposals append
straight into rank 0's memory.  Two processes on ONE GPU (CUDA IPC works between processes on the same device;
gloo carries the handle and the barrier), so the test runs on the single-GPU box; bench.py --gpus N exercises the
same code over NVLink with NCCL."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


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
    [p.start() for p in procs]
    res = sorted(q.get(timeout=240) for _ in range(2))
    [p.join(60) for p in procs]
    assert res == [(0, True, ""), (1, True, "")], res
