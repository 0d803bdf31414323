"""Multi-GPU plumbing: images are independent (lib/detect/test.py:508-513), so a batch is sharded
image-wise across ranks (one process per GPU) with NO collective on the hot path; the only exchange is
one gather of the fixed-shape per-image proposal lists at the end (NCCL over NVLink on GPUs, gloo in
the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_images(num_images: int, rank: int, world: int):
    """Contiguous block partition of image indices: rank r gets [lo, hi)."""
    per = (num_images + world - 1) // world
    lo = min(rank * per, num_images)
    return lo, min(lo + per, num_images)


def _pad_rows(t: torch.Tensor, rows: int, fill):
    """Pad the leading dimension to `rows` (shard_images hands out uneven or empty shards whenever
    num_images % world != 0; all_gather_into_tensor needs identical shapes on every rank)."""
    if t.shape[0] == rows:
        return t.contiguous()
    out = torch.full((rows,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
    out[:t.shape[0]] = t
    return out


def _coll_device(device):
    """Where an end-of-job collective runs: gloo moves host tensors (the CPU tests, and GPU tests with several processes
    on one device), NCCL device tensors."""
    return torch.device("cpu") if dist.get_backend() == "gloo" else device


def _max_rows(n: int, device) -> int:
    """Largest shard size over the ranks (one tiny all_reduce; end-of-job paths only)."""
    t = torch.tensor([int(n)], dtype=torch.int64, device=_coll_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def gather_proposals(boxes: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor):
    """all_gather of [imgs_per_rank, P, 4] f64 boxes, [imgs_per_rank, P] f32 scores and [imgs_per_rank] i32
    counts; returns rank-ordered (= image-ordered for shard_images) concatenations on every rank.  Shards may be
    uneven: every rank's block is padded to the largest shard with count-0 rows, so rank r's images start at row
    r * ceil(num_images / world) -- exactly shard_images' `lo`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return boxes, scores, counts
    world = dist.get_world_size()
    rows = _max_rows(boxes.shape[0], boxes.device)
    outs = []
    for t in (boxes, scores, counts):
        home = t.device
        t = _pad_rows(t, rows, 0).to(_coll_device(home))
        buf = torch.empty((world * rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t)
        outs.append(buf.to(home))
    return tuple(outs)


class _RawCuda:
    """A device pointer dressed up for torch.as_tensor (CUDA array interface, version 2)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerWindow:
    """`world` regions of `nbytes` bytes in the device memory of rank `root`, mapped into every other rank's process
    (libaznet_b200: azn_peer_alloc / azn_peer_open, CUDA IPC with peer access over NVLink).  Rank r appends its
    proposal lists by storing through `rank_ptr(r)` -- the last kernel of its search step does it, inside the CUDA
    graph -- so the job's exchange overlaps the search batch by batch and costs no collective kernel; `fence()` at the
    end is the only synchronisation.  This is the "gather of per-image proposal lists at the end" of test_proposals
    (one proposals.pkl written by one process, lib/detect/test.py:533-539) with the copy moved off the critical path.
    Construction is collective and all-or-nothing: if any rank cannot map the window every rank raises RuntimeError
    (the caller then keeps the all_gather route)."""

    def __init__(self, nbytes: int, device, root: int = 0):
        import ctypes as C
        from . import _lib as L
        self.rank, self.world, self.root = dist.get_rank(), dist.get_world_size(), int(root)
        self.per = (int(nbytes) + 255) // 256 * 256
        self.device = torch.device(device)
        self.base, self._owner, err = 0, self.rank == self.root, ""
        handle = [None]
        if self._owner:
            ptr, hbuf = C.c_void_p(), C.create_string_buffer(64)
            rc = L.lib().azn_peer_alloc(self.per * self.world, C.byref(ptr), hbuf)
            if rc == 0:
                self.base, handle[0] = int(ptr.value), hbuf.raw
            else:
                err = L.lib().azn_last_error().decode()
        dist.broadcast_object_list(handle, src=self.root)
        if not self._owner and handle[0] is not None:
            ptr = C.c_void_p()
            rc = L.lib().azn_peer_open(handle[0], C.byref(ptr))
            if rc == 0:
                self.base = int(ptr.value)
            else:
                err = L.lib().azn_last_error().decode()
        oks = [None] * self.world
        dist.all_gather_object(oks, (self.base != 0, err))
        if not all(o[0] for o in oks):
            self.close()
            raise RuntimeError("peer window unavailable: " + "; ".join("rank %d: %s" % (r, o[1] or "no handle") for r, o in enumerate(oks) if not o[0]))
        self._nccl = dist.get_backend() == "nccl"
        self._tok = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._tensor = None

    def rank_ptr(self, r: int | None = None) -> int:
        return self.base + (self.rank if r is None else int(r)) * self.per

    def tensor(self) -> torch.Tensor:
        """uint8 [world, per] view of the whole window (every rank can read it; rank `root` reads local memory)."""
        if self._tensor is None:
            self._tensor = torch.as_tensor(_RawCuda(self.base, self.per * self.world), device=self.device).view(self.world, self.per)
        return self._tensor

    def fence(self):
        """Every rank's stores through the window (stream-ordered before this call) are visible to every rank's work
        enqueued after it.  NCCL: one 4-byte all_reduce on the current stream (no host synchronisation)."""
        if self._nccl:
            dist.all_reduce(self._tok)
        else:
            torch.cuda.synchronize(self.device)
            dist.barrier()

    def close(self):
        from . import _lib as L
        if self.base:
            self._tensor = None
            (L.lib().azn_peer_free if self._owner else L.lib().azn_peer_close)(self.base)
            self.base = 0


class ProposalCollector:
    """Per-rank accumulation of the proposal lists of a run of batches, gathered ONCE at the end -- the way
    test_proposals appends per image and writes one proposals.pkl after the loop (lib/detect/test.py:508-539).
    `add(i, ...)` is three small device-to-device copies on the current stream (no collective per batch);
    `gather()` is the job's only exchange: one all_gather per tensor, image order = (rank, batch, image) for a
    contiguous block partition of the batches."""

    def __init__(self, n_batches: int, boxes: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor, storage: torch.Tensor | None = None):
        self.n_batches = nb = int(n_batches)
        # one byte buffer [all boxes slots | all scores slots | all counts slots]: the gather is ONE collective
        # (`storage`: a slice of a CollectorGroup's joint buffer, so that several engines still gather in one call)
        sizes = [nb * t.numel() * t.element_size() for t in (boxes, scores, counts)]
        self._buf = torch.zeros(sum(sizes), dtype=torch.uint8, device=boxes.device) if storage is None else storage
        assert self._buf.numel() == sum(sizes) and self._buf.dtype == torch.uint8
        views, off = [], 0
        for t, sz in zip((boxes, scores, counts), sizes):
            views.append(self._buf[off:off + sz].view(t.dtype).view((nb,) + tuple(t.shape)))
            off += sz
        self.boxes, self.scores, self.counts = views
        self._sizes = sizes
        self._gathered = None
        self._dst = tuple(v.data_ptr() for v in views)                            # where device_add writes
        self._remote = False
        self.state = torch.zeros(2, dtype=torch.int32, device=boxes.device)      # device-side batch counter

    def reset(self):
        self.state.zero_()

    def redirect(self, base_ptr: int):
        """device_add writes its slots at `base_ptr` (same packing as the local buffer) instead -- a region of a
        PeerWindow: the append lands in the collecting rank's memory."""
        off, dst = 0, []
        for sz in self._sizes:
            dst.append(int(base_ptr) + off)
            off += sz
        self._dst, self._remote = tuple(dst), True

    def device_add(self, boxes: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor):
        """One kernel launch (azn_collect_proposals) whose slot comes from the device-side counter: the launch is
        the same every batch, so it is captured into the CUDA graph of the search."""
        from . import _lib as L
        n_img, cap = int(boxes.shape[0]), int(boxes.shape[1])
        L.check(L.lib().azn_collect_proposals(boxes.data_ptr(), scores.data_ptr(), counts.data_ptr(), n_img, cap,
                                              self._dst[0], self._dst[1], self._dst[2],
                                              self.n_batches, self.state.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "azn_collect_proposals")

    def add(self, i: int, boxes: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor):
        if self._remote:
            raise RuntimeError("ProposalCollector.add: the collection lives in a peer window; use device_add")
        j = i % self.n_batches
        self.boxes[j].copy_(boxes, non_blocking=True)
        self.scores[j].copy_(scores, non_blocking=True)
        self.counts[j].copy_(counts, non_blocking=True)

    def gather(self, views: bool = False):
        """-> (boxes [world*n_batches*imgs, P, 4], scores [.., P], counts [..]) on every rank.  Every rank must have
        been built with the same n_batches and buffer shapes (the collective needs identical sizes; a rank with fewer
        batches leaves count-0 slots).  views=True skips the three re-packing copies and returns strided views
        [world, n_batches*imgs, ...] into the receive buffer (valid until the next gather)."""
        flat = lambda t: t.reshape((t.shape[0] * t.shape[1],) + tuple(t.shape[2:]))
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            if views:
                return tuple(flat(t).unsqueeze(0) for t in (self.boxes, self.scores, self.counts))
            return flat(self.boxes), flat(self.scores), flat(self.counts)
        world = dist.get_world_size()
        if self._gathered is None:
            self._gathered = torch.empty((world, self._buf.numel()), dtype=torch.uint8, device=self._buf.device)
        dist.all_gather_into_tensor(self._gathered.view(-1), self._buf)
        outs, off = [], 0
        for t, sz in zip((self.boxes, self.scores, self.counts), self._sizes):
            g = self._gathered[:, off:off + sz]
            if views:
                outs.append(g.view(t.dtype).view((world, t.shape[0] * t.shape[1]) + tuple(t.shape[2:])))
            else:
                outs.append(g.contiguous().view(t.dtype).view((world * t.shape[0] * t.shape[1],) + tuple(t.shape[2:])))
            off += sz
        return tuple(outs)


class CollectorGroup:
    """One ProposalCollector per engine of a multi-stream run (several batches in flight, each engine with its own
    device-side slot counter) over ONE joint byte buffer, so that the end-of-job exchange stays a single all_gather.
    `gather()` -> per collector the strided views [world, n_batches*imgs, ...] of `ProposalCollector.gather(views=True)`."""

    @staticmethod
    def nbytes(n_batches, boxes, scores, counts):
        return sum(int(n_batches) * t.numel() * t.element_size() for t in (boxes, scores, counts))

    def __init__(self, n_batches: int, outputs, peer: bool = False):
        """outputs: per engine its (out_boxes, out_scores, out_count) tensors (identical shapes).
        peer=True (multi-rank, one box): the collection lives in a PeerWindow of rank 0 and every engine's
        azn_collect_proposals appends through it; `gather()` is then a fence, not a transfer.  Raises RuntimeError
        on every rank when the window cannot be set up."""
        outputs = list(outputs)
        per = self.nbytes(n_batches, *outputs[0])
        per_al = (per + 255) // 256 * 256
        self._per, self._per_al = per, per_al
        self._buf = torch.zeros(per_al * len(outputs), dtype=torch.uint8, device=outputs[0][0].device)
        self.collectors = [ProposalCollector(n_batches, *o, storage=self._buf[k * per_al:k * per_al + per]) for k, o in enumerate(outputs)]
        self._gathered = None
        self.window = None
        if peer and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.window = PeerWindow(self._buf.numel(), self._buf.device)
            for k, c in enumerate(self.collectors):
                c.redirect(self.window.rank_ptr() + k * per_al)

    def reset(self):
        for c in self.collectors:
            c.reset()

    def gather(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return [c.gather(views=True) for c in self.collectors]
        world = dist.get_world_size()
        if self.window is not None:
            self.window.fence()                               # the lists are already in rank 0's memory
            self._gathered = self.window.tensor()
        else:
            if self._gathered is None:
                self._gathered = torch.empty((world, self._buf.numel()), dtype=torch.uint8, device=self._buf.device)
            dist.all_gather_into_tensor(self._gathered.view(-1), self._buf)
        out = []
        for k, c in enumerate(self.collectors):
            base, off, views = k * self._per_al, 0, []
            for t, sz in zip((c.boxes, c.scores, c.counts), c._sizes):
                g = self._gathered[:, base + off:base + off + sz]
                views.append(g.view(t.dtype).view((world, t.shape[0] * t.shape[1]) + tuple(t.shape[2:])))
                off += sz
            out.append(tuple(views))
        return out


def gather_detection_scores(top_scores: torch.Tensor, det_count: torch.Tensor):
    """The exchange step of the detection path: test_net's thresh[j] is global over the image set
    (lib/detect/test.py:624-631) but order-independent, so every rank all-gathers the [imgs, C, 100] f32 score
    tensor and the [imgs, C] counts of its shard and computes identical thresholds."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return top_scores, det_count
    world = dist.get_world_size()
    rows = _max_rows(top_scores.shape[0], top_scores.device)        # uneven shards: pad with empty images (count 0, -inf),
    outs = []                                                       # which never enter the thresholds
    for t, fill in ((top_scores, float("-inf")), (det_count, 0)):
        home = t.device
        t = _pad_rows(t, rows, fill).to(_coll_device(home))
        buf = torch.empty((world * rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t)
        outs.append(buf.to(home))
    return tuple(outs)
