"""Measurement pieces shared by bench.py (`extra`, `e2e_entry`, `parity` blocks) and the stand-alone tools:
BASELINE.json's other two metrics (ROI-pool HBM GB/s, NMS boxes/s: config #4), config #3 (proposals + Fast R-CNN
detection), config #2 on the default-cfg 38x63 map, and throughput through the reference's entry point
`detect.test.test_proposals`.  Everything is timed with CUDA events on torch's current stream, after warm-up, with
an L2 flush (256 MB write) between iterations of the microbenchmarks."""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from aznet_b200 import _lib, detector, engine, ops, synth  # noqa: E402

_FLUSH = {}


def flush_buffer(dev):
    if dev not in _FLUSH:
        _FLUSH[dev] = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    return _FLUSH[dev]


def time_best(fn, dev, iters=10, warm=3):
    """Best and mean of `iters` event-timed calls, L2 flushed before each."""
    flush = flush_buffer(dev)
    for _ in range(warm):
        fn()
    best, tot = 1e9, 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best, tot = min(best, ms), tot + ms
    return best, tot / iters


def hbm_write_only_gbs(dev, nbytes=1 << 30, iters=5):
    """Write-only HBM rate of a plain streaming-store kernel (azn_hbm_write_probe): what a store-dominated kernel can
    be compared with -- MEASURED_PEAKS.json's copy figure counts the read AND the write of every byte."""
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: _lib.check(_lib.lib().azn_hbm_write_probe(buf.data_ptr(), nbytes, st), "azn_hbm_write_probe")
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return nbytes / (best * 1e-3) / 1e9


def roi_pool_sweep(dev, hbm_peak, sizes=(2000, 8000, 20000), maps=((38, 63), (30, 50)), variants=("nhwc_bf16", "nhwc_f32", "nchw_f32"),
                   iters=10, C=512):
    """ROI max-pool achieved GB/s on the ALGORITHMIC bytes: map read once + 20 B per ROI + pooled rows written once
    (SURVEY 8d).  One row per (map, R, variant)."""
    rows = []
    for H, W in maps:
        feat = torch.from_numpy(synth.make_conv_maps(1, C, H, W, seed=7)).to(dev)
        nhwc32 = feat.permute(0, 2, 3, 1).contiguous()
        tens = {"nhwc_bf16": (nhwc32.to(torch.bfloat16), "NHWC", 2), "nhwc_f32": (nhwc32, "NHWC", 4), "nchw_f32": (feat, "NCHW", 4)}
        for R in sizes:
            rois = torch.from_numpy(synth.make_rois(R, 600, 1000, seed=3)).to(dev)
            for name in variants:
                f, layout, esz = tens[name]
                out = torch.empty((R, 7, 7, C) if layout == "NHWC" else (R, C, 7, 7), dtype=f.dtype, device=dev)
                best, mean = time_best(lambda: ops.roi_pool(f, rois, layout=layout, out=out), dev, iters)
                nbytes = C * H * W * esz + R * (20 + C * 49 * esz)
                gbs = nbytes / (best * 1e-3) / 1e9
                rows.append({"variant": name, "map": [H, W], "R": R, "ms": round(best, 4), "ms_mean": round(mean, 4), "bytes": nbytes,
                             "gbs": round(gbs, 1), "frac": round(gbs / hbm_peak, 4)})
    return rows


def nms_sweep(dev, sizes=(2000, 8000, 20000), threshes=(0.3, 0.7), oracle=None, match_upto=20000, iters=10):
    """Greedy NMS boxes/s (N / device time of the whole azn_nms call: sort + mask + greedy chain + compaction);
    `match` = keep list identical to the oracle of lib/utils/nms.pyx (None when not checked)."""
    rows = []
    for N in sizes:
        d_host = synth.make_dets(N, seed=3)
        d = torch.from_numpy(d_host).to(dev)
        for th in threshes:
            res = {}

            def run():
                res["k"], res["c"] = ops.nms(d, th)
            best, mean = time_best(run, dev, iters)
            kept = int(res["c"].item())
            row = {"N": N, "thresh": th, "ms": round(best, 4), "ms_mean": round(mean, 4), "kept": kept,
                   "boxes_per_s": round(N / (best * 1e-3), 0), "match": None}
            if oracle is not None and N <= match_upto:
                t0 = time.perf_counter()
                k = oracle.nms(d_host, th)
                row["cpu_oracle_s"] = round(time.perf_counter() - t0, 3)
                row["match"] = bool(k == res["k"][:kept].cpu().tolist())
            rows.append(row)
    return rows


def search_throughput(dev, head, cfg, im_h=600, im_w=1000, batch=64, steps=10, warmup=3):
    """Device-resident search images/s for one configuration (CUDA-graph replay over two rotating input sets)."""
    eng = engine.SearchEngine(head, batch, im_h, im_w, **cfg)
    fh, fw = synth.conv_shape(im_h, im_w, eng.scale)
    sets = [ops.nchw_to_nhwc_bf16(torch.from_numpy(synth.make_conv_maps(batch, head.C, fh, fw, seed=7 + 100 * s)).to(dev)) for s in range(2)]
    for i in range(warmup):
        eng.propose(sets[i % 2])
    graphs = [eng.capture(s)[0] for s in sets]
    for i in range(warmup):
        graphs[i % 2].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        graphs[i % 2].replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if int(eng.status.item()) != 0:
        raise RuntimeError("search capacity overflow")
    return {"images_per_s": round(batch / (ms * 1e-3), 1), "ms_per_step": round(ms, 4), "map": [fh, fw],
            "regions_per_image": round(float(eng.n_eval.float().mean().item()), 2)}


def detection_throughput(dev, az_head, cfg, classes=81, im_h=600, im_w=1000, batch=64, steps=5, warmup=3):
    """BASELINE config #3: search + Fast R-CNN head (COCO: 81 classes) + test_net selection + NMS 0.5, device-resident."""
    from aznet_b200.net import FRCNNHeadWeights
    cfg = dict(cfg, batch_size=10000)                               # coco.yml keeps the default SEAR.BATCH_SIZE
    eng = engine.SearchEngine(az_head, batch, im_h, im_w, **cfg)
    head = FRCNNHeadWeights(synth.make_frcnn_weights(seed=4, num_classes=classes), dev)
    det = detector.DetectEngine(head, batch, im_h, im_w, eng.cap_out, max_size=cfg["max_size"], batch_size=cfg["batch_size"])
    dset = detector.DetectionSet(batch, classes, device=dev)
    fh, fw = synth.conv_shape(im_h, im_w, eng.scale)
    sets = [ops.nchw_to_nhwc_bf16(torch.from_numpy(synth.make_conv_maps(batch, az_head.C, fh, fw, seed=7 + 100 * s)).to(dev)) for s in range(2)]

    def step(i):
        eng.propose(sets[i % 2])
        det.detect(sets[i % 2], eng.out_boxes, eng.out_count, **dset.slot(0, batch))
        dset.finish(0.5)
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"images_per_s": round(batch / (ms * 1e-3), 1), "ms_per_step": round(ms, 4), "classes": classes,
            "unique_rois_per_step": int(det.m_total.item()), "detections_kept_per_step": int(dset.keep_count.sum().item())}


class _Quiet:
    """The drivers print two lines per image (the reference's own log lines); swallow them while timing."""

    def __enter__(self):
        self.buf = io.StringIO()
        self.cm = contextlib.redirect_stdout(self.buf)
        self.cm.__enter__()
        return self

    def __exit__(self, *a):
        return self.cm.__exit__(*a)


def entry_point_throughput(dev, az_head, cfg_dict, n_images=1024, distinct=32, im_h=600, im_w=1000, zoom_fraction=0.4):
    """images/s THROUGH THE REFERENCE'S ENTRY POINT: aznet_b200.detect.test.test_proposals(net, imdb) on an in-memory
    synthetic imdb of uint8 images -- read-ahead, pinned staging, H2D of the pixels, device image blob, the VGG16
    backbone (hand-written conv kernels), the batched search, D2H of the lists, proposals.pkl.  Host wall clock around the
    call (the call returns after the pickle is written).  The backbone-only rate over the same images is reported
    beside it."""
    from aznet_b200 import backbone, net
    from aznet_b200.detect import config as C
    from aznet_b200.detect import test as T
    bw = backbone.make_vgg16_weights(seed=5)
    # a He-normal stack keeps the scale of its input, and the input here is raw mean-subtracted pixels (+-128): scale
    # conv1_1 by 1/128 (what the learned filters of a real model absorb) so that conv5_3 -- and with it the head's
    # logits -- are O(1) like the synthetic maps of the headline, instead of saturating every sigmoid
    bw["conv1_1"] = (bw["conv1_1"][0] / np.float32(128.0), bw["conv1_1"][1])
    bb = backbone.VGG16Backbone(bw, dev)
    nets = {"full": net.Net(az_head, "az", backbone=bb, name="az_vgg16"), "fc": net.Net(az_head, "az", name="az_vgg16")}
    base = synth.make_images(distinct, im_h, im_w, seed=1000)
    images = [base[i % distinct] for i in range(n_images)]
    imdb = synth.InMemoryImdb(images, num_classes=21, name="bench_mem")
    saved = (C.cfg.TEST.MAX_SIZE, C.cfg.SEAR.BATCH_SIZE, C.cfg.ROOT_DIR, dict(C.cfg.SEAR))
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        C.cfg.TEST.MAX_SIZE, C.cfg.SEAR.BATCH_SIZE, C.cfg.ROOT_DIR = cfg_dict["max_size"], cfg_dict["batch_size"], tmp
        C.cfg_set_path("bench")
        try:
            # Zoom threshold calibrated the reference's way (tools/set_thresh.py -> lib/detect/tune.py): the diagnostic
            # search in Train mode (Tz = 0) records every anchor's zoom score; Tz = the quantile that lets
            # `zoom_fraction` of the regions zoom (the conv5_3 of a random-init backbone has its own scale, so the
            # zoom bias tuned for the synthetic maps of the headline says nothing here)
            from aznet_b200.detect import tune as U
            C.cfg_set_mode("Train")
            with _Quiet():
                zs = np.concatenate([U.im_propose(nets, im)[1][:, 4] for im in base[:2]])
            tz = float(np.quantile(zs, 1.0 - zoom_fraction))
            C.cfg_set_mode("Test", tz)
            C.cfg.SEAR.NUM_PROPOSALS = cfg_dict["num_proposals"]
            with _Quiet():
                import torch.distributed as dist
                world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
                # (under torch.distributed test_proposals shards the images: every rank warms up on a full batch of its own)
                T.test_proposals(nets, synth.InMemoryImdb(images[:64 * world], num_classes=21, name="bench_warm"))    # engines, pinned rings
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                T.test_proposals(nets, imdb)
                dt = time.perf_counter() - t0
            st = dict(T.test_proposals.last_stats)
            out = {"value": round(n_images / dt, 1), "unit": "images/s", "images": n_images, "seconds": round(dt, 4),
                   "route": st["route"], "batches": st["batches"], "regions_per_image": round(st["num_eval"] / n_images, 1), "Tz": tz,
                   "host_seconds": {k: round(v, 4) for k, v in st.items() if k.startswith("host_")},
                   "h2d_bytes_per_image": st["h2d_bytes"] // n_images, "d2h_bytes_per_image": st["d2h_bytes"] // n_images,
                   "includes": "imdb read-ahead (in-memory arrays), pinned staging, H2D of uint8 pixels, azn_image_blob, VGG16 conv1_1..conv5_3, "
                               "batched adaptive search, D2H, proposals.pkl"}
        finally:
            C.cfg.TEST.MAX_SIZE, C.cfg.SEAR.BATCH_SIZE, C.cfg.ROOT_DIR = saved[:3]
            for k in ("Tz", "NUM_PROPOSALS"):
                if k in saved[3]:
                    C.cfg.SEAR[k] = saved[3][k]
                else:
                    C.cfg.SEAR.pop(k, None)
    # backbone alone on the same pixels already resident on the device (32 images per pass, like the driver)
    pix = torch.from_numpy(np.stack(base[:32])).to(dev)
    scale = engine.im_scale_for(im_h, im_w, (600,), cfg_dict["max_size"])
    for _ in range(2):
        bb.from_images(pix, scale)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        bb.from_images(pix, scale)
    e1.record()
    torch.cuda.synchronize()
    out["backbone_only_images_per_s"] = round(64 / (e0.elapsed_time(e1) * 1e-3), 1)
    return out
