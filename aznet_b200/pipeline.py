"""Host-facing batched proposal call: conv5_3 maps as f32 NCHW host arrays in (what the reference's 'fc'
net is handed, lib/detect/test.py:228-236), proposal lists in pinned host memory out.

The reference uploads the map and downloads the head outputs once per level per image
(pycaffe.py:90,95).  Here a batch crosses PCIe once in each direction, and the upload of batch i+1
(copy stream) overlaps the search of batch i (compute stream): `submit()` returns a ticket immediately,
`result(ticket)` blocks until that batch's proposals are on the host.

PCIe bounds this call (196.6 MB of f32 maps per 64-image batch against a 0.8 ms search), and the engine stores maps
as bf16 anyway, so the batch can be NARROWED ON THE HOST before it crosses the link (`host_narrow`): the library's
worker threads round image chunk k to bf16 into pinned memory (`azn_host_f32_to_bf16`, the same rounding as the device
conversion: identical bits in HBM) while the copy engine uploads chunk k-1.  Whether that wins depends on the host
(cores per rank, memory bandwidth), so "auto" times the routes on the first batch and keeps the fastest one.

Neither resource is busy all the time on either route -- the plain upload leaves the host cores idle, the narrowed one
leaves the link waiting for the cores -- so the third route is a SPLIT (`host_narrow="split"`, and one of the candidates
of "auto"): the first `raw_images` images of the batch cross the link as f32 while the cores narrow the others, piece
by piece, and each narrowed chunk follows the raw piece that was uploading while it was rounded.  The device converts
the two parts with the two layout kernels into the same bf16 NHWC batch, so the search sees identical bits.
"""
from __future__ import annotations

import os
import time

import torch

from . import ops
from .engine import SearchEngine


def default_host_threads():
    """Worker threads for the host-side narrowing: this rank's share of the cores it may run on, at most 16."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    return max(1, min(16, n // local))


class _Slot:
    def __init__(self, eng: SearchEngine, shape, dev, dtype=torch.float32, narrow=False):
        self.stage = torch.empty(shape, dtype=dtype, device=dev)
        # host-narrowed route: pinned bf16 staging (the DMA source) + its device twin, both NCHW like the input
        self.pin16 = torch.empty(shape, dtype=torch.bfloat16).pin_memory() if narrow else None
        self.stage16 = torch.empty(shape, dtype=torch.bfloat16, device=dev) if narrow else None
        self.boxes = torch.empty(eng.out_boxes.shape, dtype=torch.float64).pin_memory()
        self.scores = torch.empty(eng.out_scores.shape, dtype=torch.float32).pin_memory()
        self.count = torch.empty(eng.out_count.shape, dtype=torch.int32).pin_memory()
        self.n_eval = torch.empty(eng.n_eval.shape, dtype=torch.int32).pin_memory()
        self.status = torch.empty(1, dtype=torch.int32).pin_memory()
        self.h2d_done = torch.cuda.Event()
        self.consumed = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.busy = False
        self.graph = None


class ProposalPipeline:
    def __init__(self, eng: SearchEngine, map_shape, depth: int = 2, after_search=None, use_graph: bool = True,
                 layout: str = "nchw_f32", host_narrow: str = "off", host_threads: int = 0, narrow_chunks: int = 8,
                 raw_fraction: float = 1.0 / 3.0):
        """map_shape = (n_img, C, H, W) of the batches; after_search: optional callable run on the compute stream right
        after the search (e.g. the NCCL gather of a multi-GPU run); use_graph: replay the layout conversion + level
        loop of every slot from a CUDA graph.
        layout "nchw_f32": host batches are f32 NCHW, what the reference's 'fc' net is handed (pycaffe blobs);
        layout "nhwc_bf16": host batches are already in the engine's own storage format, bf16 [n, H, W, C] -- half the
        PCIe bytes and no conversion kernel, for callers that keep their maps that way (e.g. downloaded from
        aznet_b200.backbone).
        host_narrow (layout "nchw_f32" only): "off" uploads the f32 batch as it is; "on" rounds it to bf16 on the host
        first, `narrow_chunks` image chunks pipelined against their uploads, on `host_threads` worker threads (0: this
        rank's share of the cores); "split" uploads the first `raw_fraction` of the images as f32 while the others are
        narrowed; "auto" times the plain upload, the narrowed route and a few splits on the first submitted batch and
        keeps the fastest (`self.narrow`, `self.raw_images`, `self.narrow_timing`)."""
        assert layout in ("nchw_f32", "nhwc_bf16")
        assert host_narrow in ("off", "on", "auto", "split")
        self.eng, self.dev, self.layout = eng, eng.dev, layout
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        n, c, h, w = map_shape
        want_narrow = layout == "nchw_f32" and host_narrow != "off"
        self.narrow = True if (want_narrow and host_narrow in ("on", "split")) else (None if want_narrow else False)   # None: undecided
        # images of a batch that cross the link as f32 (the rest is narrowed on the host): n = plain upload, 0 = all narrowed
        self.raw_images = n if not want_narrow else (0 if host_narrow == "on" else (max(1, min(n - 1, round(n * raw_fraction))) if host_narrow == "split" and n > 1 else 0))
        self._n_img = n
        self.host_threads = host_threads if host_threads > 0 else default_host_threads()
        self.narrow_chunks = max(1, min(narrow_chunks, n))
        self.narrow_timing = None
        if layout == "nchw_f32":
            self.slots = [_Slot(eng, map_shape, self.dev, narrow=want_narrow) for _ in range(depth)]
            self.nhwc = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=self.dev)
        else:
            self.slots = [_Slot(eng, (n, h, w, c), self.dev, torch.bfloat16) for _ in range(depth)]
            self.nhwc = None
        self.after_search = after_search
        self.use_graph = use_graph
        self.launches_per_submit = 0
        self._i = 0
        self._map_elems = int(n * c * h * w)
        s = self.slots[0]
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in (s.boxes, s.scores, s.count, s.n_eval, s.status))

    @property
    def h2d_bytes(self):
        """Bytes that cross PCIe per submitted batch on the route in use."""
        if self.layout == "nhwc_bf16":
            return self._map_elems * 2
        if not self.narrow:
            return self._map_elems * 4
        per = self._map_elems // self._n_img
        return per * (4 * self.raw_images + 2 * (self._n_img - self.raw_images))

    def _upload_narrowed(self, slot: _Slot, host_maps: torch.Tensor, raw: int | None = None):
        """Images [0, raw) cross the link as f32, piece by piece; image chunk k of the others is rounded to bf16 on the
        host while raw piece k (and narrowed chunk k-1) upload.  raw = 0: everything is narrowed."""
        n = host_maps.shape[0]
        m = self.raw_images if raw is None else raw
        k = self.narrow_chunks
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot.consumed)
            for i in range(k):
                ra, rb = m * i // k, m * (i + 1) // k
                if rb > ra:
                    slot.stage[ra:rb].copy_(host_maps[ra:rb], non_blocking=True)
                a, b = m + (n - m) * i // k, m + (n - m) * (i + 1) // k
                if b > a:
                    ops.host_f32_to_bf16(host_maps[a:b], slot.pin16[a:b], self.host_threads)
                    slot.stage16[a:b].copy_(slot.pin16[a:b], non_blocking=True)
            slot.h2d_done.record(self.copy_stream)

    def _upload_plain(self, slot: _Slot, host_maps: torch.Tensor):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot.consumed)          # the previous user of this staging buffer
            slot.stage.copy_(host_maps, non_blocking=True)
            slot.h2d_done.record(self.copy_stream)

    def _decide(self, slot: _Slot, host_maps: torch.Tensor):
        """host_narrow="auto": host time from call to upload complete, best of three, of every route on this batch: the
        plain upload, everything narrowed, and splits with 1/4, 3/8 and 1/2 of the images uploaded raw."""
        n = self._n_img
        routes = [("f32_upload", None), ("host_bf16_then_upload", 0)]
        for num, den in ((1, 4), (3, 8), (1, 2)):
            m = n * num // den
            if 0 < m < n and all(m != r[1] for r in routes):
                routes.append(("split_%d_of_%d_raw" % (m, n), m))
        best = {}

        def measure(name, m):
            ts = []
            for _ in range(3):
                torch.cuda.synchronize(self.dev)
                t0 = time.perf_counter()
                if m is None:
                    self._upload_plain(slot, host_maps)
                else:
                    self._upload_narrowed(slot, host_maps, m)
                slot.h2d_done.synchronize()
                ts.append(time.perf_counter() - t0)
            best[name] = min(ts)

        for name, m in routes:
            measure(name, m)
        chosen = min(best, key=best.get)
        m = dict(routes)[chosen]
        step = max(1, n // 16)
        for _ in range(6):                                   # a split won: walk a sixteenth of the batch at a time while it improves
            if not m:
                break
            tried = False
            for mm in (m - step, m + step):
                name = "split_%d_of_%d_raw" % (mm, n)
                if 0 < mm < n and name not in best:
                    routes.append((name, mm))
                    measure(name, mm)
                    tried = True
            chosen = min(best, key=best.get)
            if dict(routes)[chosen] == m or not tried:
                break
            m = dict(routes)[chosen]
        self.narrow = m is not None
        self.raw_images = n if m is None else m
        self.narrow_timing = {k: round(v * 1e3, 3) for k, v in best.items()}
        self.narrow_timing.update(unit="ms per batch", host_threads=self.host_threads, chunks=self.narrow_chunks, chosen=chosen)

    def _convert(self, slot: _Slot):
        """Staged NCHW batch (f32 images [0, raw), bf16 images [raw, n)) -> the engine's bf16 NHWC batch."""
        m = self.raw_images if self.narrow else self._n_img
        if m > 0:
            ops.nchw_to_nhwc_bf16(slot.stage[:m], out=self.nhwc[:m])
            self.eng.launches += 1
        if m < self._n_img:
            ops.nchw_bf16_to_nhwc_bf16(slot.stage16[m:], out=self.nhwc[m:])
            self.eng.launches += 1

    def submit(self, host_maps: torch.Tensor) -> _Slot:
        """host_maps: CPU tensor in the pipeline's layout (pinned for an asynchronous copy)."""
        slot = self.slots[self._i % len(self.slots)]
        self._i += 1
        if slot.busy:
            raise RuntimeError("pipeline slot still in flight: call result() on the oldest ticket first")
        slot.busy = True
        compute = torch.cuda.current_stream(self.dev)
        if self.narrow is None:
            self._decide(slot, host_maps)
        if self.narrow:
            self._upload_narrowed(slot, host_maps)
        else:
            self._upload_plain(slot, host_maps)
        compute.wait_event(slot.h2d_done)
        direct = self.layout == "nhwc_bf16"
        if self.use_graph:
            if slot.graph is None:
                def pre(slot=slot):
                    self._convert(slot)
                slot.graph, self.launches_per_submit = self.eng.capture(slot.stage if direct else self.nhwc, pre=None if direct else pre)
            slot.graph.replay()
            self.eng.launches += self.launches_per_submit
        elif direct:
            self.eng.propose(slot.stage)
        else:
            self._convert(slot)
            self.eng.propose(self.nhwc)
        slot.consumed.record(compute)
        if self.after_search is not None:
            self.after_search()
        e = self.eng
        for dst, src in ((slot.boxes, e.out_boxes), (slot.scores, e.out_scores), (slot.count, e.out_count),
                         (slot.n_eval, e.n_eval), (slot.status, e.status)):
            dst.copy_(src, non_blocking=True)
        slot.done.record(compute)
        return slot

    def result(self, slot: _Slot):
        """Blocks until the ticket's batch is on the host; returns (boxes [n,P,4] f64, scores [n,P] f32,
        counts [n] i32, regions evaluated [n] i32) as pinned host tensors (valid until the slot is reused)."""
        slot.done.synchronize()
        slot.busy = False
        if int(slot.status[0]) != 0:
            raise RuntimeError("search capacity overflow on device (status %d)" % int(slot.status[0]))
        return slot.boxes, slot.scores, slot.count, slot.n_eval
