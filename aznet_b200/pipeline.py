"""Host-facing batched proposal call: conv5_3 maps as f32 NCHW host arrays in (what the reference's 'fc'
net is handed, lib/detect/test.py:228-236), proposal lists in pinned host memory out.

The reference uploads the map and downloads the head outputs once per level per image
(pycaffe.py:90,95).  Here a batch crosses PCIe once in each direction, and the upload of batch i+1
(copy stream) overlaps the search of batch i (compute stream): `submit()` returns a ticket immediately,
`result(ticket)` blocks until that batch's proposals are on the host.
"""
from __future__ import annotations

import torch

from . import ops
from .engine import SearchEngine


class _Slot:
    def __init__(self, eng: SearchEngine, shape, dev, dtype=torch.float32):
        self.stage = torch.empty(shape, dtype=dtype, device=dev)
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
                 layout: str = "nchw_f32"):
        """map_shape = (n_img, C, H, W) of the batches; after_search: optional callable run on the compute stream right
        after the search (e.g. the NCCL gather of a multi-GPU run); use_graph: replay the layout conversion + level
        loop of every slot from a CUDA graph.
        layout "nchw_f32": host batches are f32 NCHW, what the reference's 'fc' net is handed (pycaffe blobs);
        layout "nhwc_bf16": host batches are already in the engine's own storage format, bf16 [n, H, W, C] -- half the
        PCIe bytes and no conversion kernel, for callers that keep their maps that way (e.g. downloaded from
        aznet_b200.backbone)."""
        assert layout in ("nchw_f32", "nhwc_bf16")
        self.eng, self.dev, self.layout = eng, eng.dev, layout
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        n, c, h, w = map_shape
        if layout == "nchw_f32":
            self.slots = [_Slot(eng, map_shape, self.dev) for _ in range(depth)]
            self.nhwc = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=self.dev)
        else:
            self.slots = [_Slot(eng, (n, h, w, c), self.dev, torch.bfloat16) for _ in range(depth)]
            self.nhwc = None
        self.after_search = after_search
        self.use_graph = use_graph
        self.launches_per_submit = 0
        self._i = 0
        self.h2d_bytes = int(n * c * h * w * (4 if layout == "nchw_f32" else 2))
        s = self.slots[0]
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in (s.boxes, s.scores, s.count, s.n_eval, s.status))

    def submit(self, host_maps: torch.Tensor) -> _Slot:
        """host_maps: CPU tensor in the pipeline's layout (pinned for an asynchronous copy)."""
        slot = self.slots[self._i % len(self.slots)]
        self._i += 1
        if slot.busy:
            raise RuntimeError("pipeline slot still in flight: call result() on the oldest ticket first")
        slot.busy = True
        compute = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot.consumed)          # the previous user of this staging buffer
            slot.stage.copy_(host_maps, non_blocking=True)
            slot.h2d_done.record(self.copy_stream)
        compute.wait_event(slot.h2d_done)
        direct = self.layout == "nhwc_bf16"
        if self.use_graph:
            if slot.graph is None:
                def pre(stage=slot.stage):
                    ops.nchw_to_nhwc_bf16(stage, out=self.nhwc)
                    self.eng.launches += 1
                slot.graph, self.launches_per_submit = self.eng.capture(slot.stage if direct else self.nhwc, pre=None if direct else pre)
            slot.graph.replay()
            self.eng.launches += self.launches_per_submit
        elif direct:
            self.eng.propose(slot.stage)
        else:
            ops.nchw_to_nhwc_bf16(slot.stage, out=self.nhwc)
            self.eng.launches += 1
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
