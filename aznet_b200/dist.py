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


def gather_proposals(boxes: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor):
    """all_gather of [imgs_per_rank, P, 4] f64 boxes, [imgs_per_rank, P] f32 scores and [imgs_per_rank] i32
    counts; returns rank-ordered (= image-ordered for shard_images) concatenations on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return boxes, scores, counts
    world = dist.get_world_size()
    outs = []
    for t in (boxes, scores, counts):
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t.contiguous())
        outs.append(buf)
    return tuple(outs)


def gather_detection_scores(top_scores: torch.Tensor, det_count: torch.Tensor):
    """The exchange step of the detection path: test_net's thresh[j] is global over the image set
    (lib/detect/test.py:624-631) but order-independent, so every rank all-gathers the [imgs, C, 100] f32 score
    tensor and the [imgs, C] counts of its shard (equal shard sizes) and computes identical thresholds."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return top_scores, det_count
    world = dist.get_world_size()
    outs = []
    for t in (top_scores, det_count):
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t.contiguous())
        outs.append(buf)
    return tuple(outs)
