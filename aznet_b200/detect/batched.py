"""Batched, read-ahead execution of the dataset drivers test_proposals / test_net / test_net_shared
(lib/detect/test.py:486-539, 541-668, 670-780).

The reference walks the image database one image at a time: cv2.imread, a host-side mean-subtract + cv2.resize
per search level (test.py:208), one net.forward per level, results appended to a Python list.  Images are
independent (nothing crosses images inside im_propose), so the drivers here

  * read ahead with a thread pool (cv2.imread releases the GIL) -- or take arrays from an optional
    `imdb.image_at(i)` -- and group images of the same shape into batches of up to BATCH images, staged in pinned
    host memory by the reader threads;
  * upload the uint8 pixels (3 B/pixel instead of a 12 B/pixel f32 blob) on a copy stream while the previous
    batch computes, build the network input on the device (azn_image_blob: no host cv2), run the backbone in
    chunks and the batched SearchEngine / DetectEngine over the whole batch;
  * fetch each batch's results while the next one runs.

Results are written back by image index, so proposals.pkl / detections.pkl keep the reference's order.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .config import cfg
from .. import ops
from ..detector import DetectEngine
from ..engine import SearchEngine, im_scale_for

BATCH = 64            # images per search / detection batch (BASELINE config #2's batch)
BACKBONE_CHUNK = int(os.environ.get("AZN_BACKBONE_CHUNK", "32"))   # images per backbone pass (activations of conv1_x: ~100 MB per image at 480x800; measured 3572 / 3699 / 3528 images/s at 16 / 32 / 64: conv5_x fills its second wave of tiles at 32, conv1_2 slows down at 64)
_SIZES = (1, 2, 4, 8, 16, 32, 64)


def _pad_size(n):
    """Engines exist for a few batch sizes only; a short batch is padded with empty images."""
    for s in _SIZES:
        if n <= s:
            return s
    return n


class Batch:
    """idx: image indices (database order within the batch); images: pinned uint8 [n, H, W, 3]."""
    __slots__ = ("idx", "images", "shape", "_ring", "_slot")

    def __init__(self, idx, images, shape, ring, slot):
        self.idx, self.images, self.shape, self._ring, self._slot = idx, images, shape, ring, slot

    def uploaded(self, event):
        """The consumer records `event` after the H2D copy of `images`; the staging buffer is reused after it."""
        self._ring[self._slot][1] = event


# Pinned staging buffers are page-locked allocations (cudaHostAlloc: tens of milliseconds for a 115 MB batch), so they
# are kept for the life of the process: a ring of 3 per (batch, image shape), a few shapes at a time.
_PINNED_RINGS = OrderedDict()


class ImageFeeder:
    """Iterates over same-shape batches of the images `indices` of an imdb, reading ahead `window` images."""

    def __init__(self, imdb, indices, batch=BATCH, workers=None, window=None):
        self.imdb, self.indices, self.batch = imdb, list(indices), int(batch)
        self.workers = workers or min(32, os.cpu_count() or 4)
        self.window = window or 3 * self.batch
        self._loader = imdb.image_at if hasattr(imdb, "image_at") else self._imread

    def _imread(self, i):
        import cv2
        im = cv2.imread(self.imdb.image_path_at(i))
        if im is None:
            raise IOError("cannot read image %r" % (self.imdb.image_path_at(i),))
        return im

    def _stage(self, pool, shape, items):
        """Copy a group's images into a pinned [batch, H, W, 3] buffer (ring of 3 per shape) with the reader threads."""
        key = (self.batch, shape)
        ring = _PINNED_RINGS.get(key)
        if ring is None:
            ring = _PINNED_RINGS[key] = {"next": 0}
            if len(_PINNED_RINGS) > 6:                                # bound pinned memory across many distinct shapes
                _PINNED_RINGS.popitem(last=False)
        else:
            _PINNED_RINGS.move_to_end(key)
        slot = ring["next"] % 3
        ring["next"] += 1
        if slot not in ring:
            # a database with more than two batches of this shape will cycle through all three buffers: page-lock them
            # together, once (the first batch of a shape pays ~0.1 s here instead of the third)
            for k in ((0, 1, 2) if ring["next"] > 1 or len(items) == self.batch else (slot,)):
                if k not in ring:
                    ring[k] = [torch.empty((self.batch, shape[0], shape[1], 3), dtype=torch.uint8).pin_memory(), None]
        buf, ev = ring[slot]
        if ev is not None:
            ev.synchronize()                                          # its last upload has left the buffer
        host = buf.numpy()
        list(pool.map(lambda ki: np.copyto(host[ki[0]], ki[1][1]), enumerate(items)))
        return Batch([i for i, _ in items], buf[:len(items)], shape, ring, slot)

    def __iter__(self):
        groups = OrderedDict()
        with ThreadPoolExecutor(self.workers) as pool:
            futures, nxt = [], 0
            n = len(self.indices)
            done = 0
            while done < n:
                while nxt < n and len(futures) < self.window:
                    futures.append((self.indices[nxt], pool.submit(self._loader, self.indices[nxt])))
                    nxt += 1
                i, f = futures.pop(0)
                im = f.result()
                done += 1
                if im.ndim != 3 or im.shape[2] != 3 or im.dtype != np.uint8:
                    raise ValueError("image %d: expected uint8 HxWx3 (BGR), got %s %s" % (i, im.dtype, im.shape))
                g = groups.setdefault(im.shape[:2], [])
                g.append((i, im))
                if len(g) == self.batch:
                    yield self._stage(pool, im.shape[:2], groups.pop(im.shape[:2]))
            for shape, g in groups.items():
                yield self._stage(pool, shape, g)


class _EngineCache:
    """A few engines (each holds its pooled-row and activation buffers, ~2 GB at 64 images of 600x1000) by key, LRU."""

    def __init__(self, cap=6):
        self.cap, self.d = cap, OrderedDict()

    def get(self, key, make):
        e = self.d.get(key)
        if e is None:
            while len(self.d) >= self.cap:
                self.d.popitem(last=False)
            e = self.d[key] = make()
        else:
            self.d.move_to_end(key)
        return e


_SEARCH, _DETECT = _EngineCache(), _EngineCache()


def search_engine(az_net, im_shape, n_img, num_proposals=None):
    fixed = bool(cfg.SEAR.FIXED_PROPOSAL_NUM) or num_proposals is not None
    npr = num_proposals if num_proposals is not None else cfg.SEAR.NUM_PROPOSALS
    key = (id(az_net.head), n_img, int(im_shape[0]), int(im_shape[1]), tuple(cfg.TEST.SCALES), cfg.TEST.MAX_SIZE, cfg.SEAR.MIN_SIDE,
           float(cfg.SEAR.Tz), float(cfg.SEAR.Tc), fixed, npr, cfg.SEAR.BATCH_SIZE, float(cfg.DEDUP_BOXES), float(cfg.EPS))
    return _SEARCH.get(key, lambda: SearchEngine(
        az_net.head, n_img, im_shape[0], im_shape[1], scales=tuple(cfg.TEST.SCALES), max_size=cfg.TEST.MAX_SIZE,
        min_side=cfg.SEAR.MIN_SIDE, tz=float(cfg.SEAR.Tz), tc=float(cfg.SEAR.Tc), fixed_num=fixed, num_proposals=npr,
        batch_size=cfg.SEAR.BATCH_SIZE, dedup=float(cfg.DEDUP_BOXES), eps=float(cfg.EPS), spatial_scale=az_net.spatial_scale))


def detect_engine(fr_net, im_shape, n_img, cap):
    key = (id(fr_net.head), n_img, int(im_shape[0]), int(im_shape[1]), int(cap), tuple(cfg.TEST.SCALES), cfg.TEST.MAX_SIZE,
           cfg.SEAR.BATCH_SIZE, float(cfg.DEDUP_BOXES), float(cfg.EPS))
    return _DETECT.get(key, lambda: DetectEngine(
        fr_net.head, n_img, im_shape[0], im_shape[1], cap, scales=tuple(cfg.TEST.SCALES), max_size=cfg.TEST.MAX_SIZE,
        batch_size=cfg.SEAR.BATCH_SIZE, dedup=float(cfg.DEDUP_BOXES), eps=float(cfg.EPS), spatial_scale=fr_net.spatial_scale))


class DeviceBatch:
    """One batch on the device: uploaded pixels -> network input -> backbone maps.  `maps` is the bf16 NHWC conv5_3
    batch [n_pad, fh, fw, C] (rows past n are zero: padding images), or a dict of taps for the skip-layer head."""

    def __init__(self, full_net, batch: Batch, copy_stream, taps=("conv5_3",)):
        dev = full_net.dev
        n, (H, W) = len(batch.idx), batch.shape
        self.n, self.n_pad, self.shape, self.idx = n, _pad_size(n), (H, W), batch.idx
        compute = torch.cuda.current_stream(dev)
        with torch.cuda.stream(copy_stream):
            pix = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
            pix.copy_(batch.images, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        batch.uploaded(ev)
        compute.wait_event(ev)
        pix.record_stream(compute)
        self.h2d_bytes = pix.numel()
        bb = full_net.backbone
        scale = im_scale_for(H, W, tuple(cfg.TEST.SCALES), cfg.TEST.MAX_SIZE)
        taps = tuple(taps)
        outs = {t: None for t in taps}
        for lo in range(0, n, BACKBONE_CHUNK):
            hi = min(lo + BACKBONE_CHUNK, n)
            blob = ops.image_blob(pix[lo:hi], scale, bb.pixel_means, bb.cpad_in)
            res = bb.run_padded(blob, taps=taps)
            for t in taps:
                m = res[t]
                if outs[t] is None:
                    outs[t] = torch.zeros((self.n_pad,) + tuple(m.shape[1:]), dtype=m.dtype, device=dev)
                outs[t][lo:hi] = m
        self.launches = ((n + BACKBONE_CHUNK - 1) // BACKBONE_CHUNK) * (bb.launches_per_call + len(taps))
        self.maps = outs["conv5_3"] if taps == ("conv5_3",) else outs


_PINNED_POOL = {}


class Fetch:
    """Device tensors -> pinned host copies on the current stream; `get()` waits for them and hands out arrays that
    stay valid (copies): the pinned buffers go back to a small pool."""

    def __init__(self, **tensors):
        self.host = {}
        for k, t in tensors.items():
            key = (tuple(t.shape), t.dtype)
            free = _PINNED_POOL.setdefault(key, [])
            h = free.pop() if free else torch.empty(t.shape, dtype=t.dtype).pin_memory()
            h.copy_(t, non_blocking=True)
            self.host[k] = h
        self.bytes = sum(h.numel() * h.element_size() for h in self.host.values())
        self.done = torch.cuda.Event()
        self.done.record()

    def get(self):
        self.done.synchronize()
        out = {k: v.numpy().copy() for k, v in self.host.items()}
        for v in self.host.values():
            free = _PINNED_POOL[(tuple(v.shape), v.dtype)]
            if len(free) < 4:
                free.append(v)
        self.host = {}
        return out
