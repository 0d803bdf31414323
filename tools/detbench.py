#!/usr/bin/env python
"""BASELINE config #3: AZ-Net proposals + Fast R-CNN detection head (COCO: 81 classes) + test_net selection + NMS
for batches of 64 synthetic 600x1000 images on one B200, everything device-resident (shared conv5_3 maps).

A step = search (300 proposals per image) -> DetectEngine (dedup, staged ROI pool, fc6/fc7/cls|bbox, per-class
top-100 + decode) -> set-wide thresholds + filter + NMS of the 64 x 80 problems.  CUDA-event timing per stage,
weights + pooled rows exceed L2.  Prints one JSON line; optional CPU leg = the oracle port on the host cores.

    python tools/detbench.py [--steps 10] [--warmup 3] [--classes 81] [--cpu-images 2]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

IM_H, IM_W, BATCH = 600, 1000, 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--classes", type=int, default=81)
    ap.add_argument("--pool", default="staged", choices=("staged", "direct", "per_roi"), help="ROI-pool kernel of the detection step (A/B)")
    ap.add_argument("--cpu-images", type=int, default=2)
    args = ap.parse_args()
    import torch
    from aznet_b200 import _lib, detector, engine, ops, synth
    from aznet_b200.net import FRCNNHeadWeights
    import bench as B

    _lib.build()
    _lib.require_device()
    dev = torch.device("cuda:0")
    detector.POOL_KERNEL = args.pool
    torch.cuda.set_device(dev)
    cfg = dict(B.CFG, batch_size=10000)                       # coco.yml: default SEAR.BATCH_SIZE
    azw = synth.make_az_weights(seed=3, zoom_bias=B.ZOOM_BIAS)
    frw = synth.make_frcnn_weights(seed=4, num_classes=args.classes)
    eng = engine.SearchEngine(engine.AZHeadWeights(azw, dev), BATCH, IM_H, IM_W, **cfg)
    head = FRCNNHeadWeights(frw, dev)
    det = detector.DetectEngine(head, BATCH, IM_H, IM_W, eng.cap_out, max_size=cfg["max_size"], batch_size=cfg["batch_size"])
    dset = detector.DetectionSet(BATCH, args.classes, device=dev)
    fh, fw = synth.conv_shape(IM_H, IM_W, eng.scale)
    maps = [ops.nchw_to_nhwc_bf16(torch.from_numpy(synth.make_conv_maps(BATCH, 512, fh, fw, seed=7 + 100 * s)).to(dev))
            for s in range(2)]
    stages = ["search", "rois", "roi_pool", "fc6", "fc7", "cls_bbox", "select", "finish"]
    acc = {k: 0.0 for k in stages}
    m_rows = []

    def step(i, timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
        conv = maps[i % 2]
        hd = det.head
        mc = det.n_img * det.cap
        ev[0].record()
        eng.propose(conv)
        ev[1].record()
        det.prepare(eng.out_boxes, eng.out_count)
        ev[2].record()
        pool = ops.roi_pool(conv, det.rois, hd.pooled, det.spatial_scale, layout="NHWC", n_rois=det.m_total,
                            out=det.pool5.view(mc, hd.pooled, hd.pooled, hd.C), **detector.pool_kwargs())
        ev[3].record()
        ops.fc_forward(pool.view(mc, -1), hd.w6, hd.b6, _lib.ACT_RELU, m_live=det.m_total, out=det.h6)
        ev[4].record()
        ops.fc_forward(det.h6, hd.w7, hd.b7, _lib.ACT_RELU, m_live=det.m_total, out=det.h7)
        ev[5].record()
        ops.fc_forward(det.h7, hd.wo, hd.bo, _lib.ACT_SOFTMAX_BBOX, det.C, m_live=det.m_total, out=det.out[:, :det.n_out])
        ev[6].record()
        det.select(**dset.slot(0, BATCH))
        ev[7].record()
        dset.finish(0.5)
        ev[8].record()
        if timed:
            torch.cuda.synchronize()
            for k, name in enumerate(stages):
                acc[name] += ev[k].elapsed_time(ev[k + 1])
            m_rows.append(int(det.m_total.item()))

    for i in range(max(args.warmup, 3)):
        step(i, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    for i in range(args.steps):                               # second pass: per-stage events (sync per step)
        step(i, True)
    per = {k: round(v / args.steps, 4) for k, v in acc.items()}
    m = float(np.mean(m_rows))
    hbm_peak, tf_peak, which = B.peaks()
    k6, n6 = head.w6.shape[1], head.w6.shape[0]
    fc6_tf = 2.0 * m * n6 * k6 / (per["fc6"] * 1e-3) / 1e12
    pool_gbs = (BATCH * fh * fw * 512 * 2 + m * (20 + k6 * 2)) / (per["roi_pool"] * 1e-3) / 1e9
    kept = int(dset.keep_count.sum().item())
    line = {
        "metric": "AZ proposals + Fast R-CNN detection images/sec", "value": BATCH / (ms / 1e3), "unit": "images/s",
        "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE config #3: AZ-Net VGG16 COCO config (coco.yml: MAX_SIZE 800), batch of 64 synthetic "
                               "600x1000 images, 300 proposals/image, Fast R-CNN head with %d classes, test_net selection, "
                               "NMS 0.5; shared cached conv5_3" % args.classes,
                   "unique_rois_per_step": m, "detections_kept_per_step": kept},
        "per_stage_ms": per,
        "roofline": {"kernel": "fc_gemm_kernel<256,2> fc6 25088->4096, M=%d" % int(m), "bound": "tensor", "achieved": fc6_tf,
                     "peak": tf_peak, "unit": "TFLOP/s", "frac": fc6_tf / tf_peak, "peak_source": which},
        "roi_pool": {"gbs": pool_gbs, "frac_of_hbm": pool_gbs / hbm_peak, "bytes_model": "maps once + pooled rows written"},
        "gpu_launches_per_step": (eng.launches + det.launches) // max(2 * args.steps + max(args.warmup, 3), 1),
    }
    if args.cpu_images > 0:
        from oracle import az_oracle as O
        threads = os.cpu_count() or 1
        ocfg = O.OracleCfg(TEST_MAX_SIZE=cfg["max_size"], Tz=cfg["tz"], NUM_PROPOSALS=cfg["num_proposals"], BATCH_SIZE=cfg["batch_size"])
        conv = synth.make_conv_maps(args.cpu_images, 512, fh, fw, seed=7)
        aznet = O.OracleNet(azw, "az", cfg=ocfg, threads=threads)
        frnet = O.OracleNet(frw, "frcnn", cfg=ocfg, threads=threads)
        t0 = time.perf_counter()
        per_image = []
        for i in range(args.cpu_images):
            c = {"conv5_3": conv[i:i + 1]}
            Y = O.im_propose({"full": aznet, "fc": aznet}, (IM_H, IM_W, 3), ocfg, conv=c)
            s, p, _ = O.frcnn_forward({"full": frnet, "fc": frnet}, (IM_H, IM_W, 3), Y, args.classes, c, ocfg)
            per_image.append((s, p))
        ab, _ = O.test_net_select(per_image, args.classes)
        O.apply_nms(ab, 0.5)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": args.cpu_images / dt, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": "%d images in %.1f s (oracle port: search + Fast R-CNN head + selection + NMS)" % (args.cpu_images, dt)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
