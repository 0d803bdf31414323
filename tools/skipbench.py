#!/usr/bin/env python
"""Skip-layer detector head at full size: models/COCO/VGG16_skip/frcnn/test_fc.prototxt with
experiments/cfgs/voc_skip.yml (MAX_SIZE 800, DEDUP_BOXES 0.5) on a batch of 64 synthetic 600x1000 images, 300
proposals each: ROI pools over conv3_3 [120x200x256], conv4_3 [60x100x512], conv5_3 [30x50x512], GRN + concat + x1000,
conv_pool5 (1280 -> 512 over 49 positions per ROI), fc6 / fc7 / cls_score | bbox_pred, test_net selection.
Per-stage device times (CUDA events) with the algorithmic GB/s / TFLOP/s of each stage.  One JSON line.

    python tools/skipbench.py [--steps 5] [--warmup 3] [--classes 21]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

IM_H, IM_W, BATCH, NPROP = 600, 1000, 64, 300


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--classes", type=int, default=21)
    ap.add_argument("--unfused", action="store_true", help="separate ROI pools + azn_grn_concat_forward instead of the fused epilogue")
    args = ap.parse_args()
    import torch
    from aznet_b200 import _lib, detector, engine, ops, synth
    from aznet_b200.net import FRCNNSkipHeadWeights
    import bench as B

    _lib.build()
    _lib.require_device()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    head = FRCNNSkipHeadWeights(synth.make_frcnn_skip_weights(seed=4, num_classes=args.classes), dev)
    det = detector.DetectEngine(head, BATCH, IM_H, IM_W, NPROP, max_size=800, batch_size=1000, dedup=0.5)
    dset = detector.DetectionSet(BATCH, args.classes, device=dev)
    scale = engine.im_scale_for(IM_H, IM_W, (600,), 800)
    hs, ws = int(round(IM_H * scale)), int(round(IM_W * scale))
    shapes, h, w = {}, hs, ws
    for name, c, pools in (("conv3_3", 256, 2), ("conv4_3", 512, 1), ("conv5_3", 512, 1)):
        for _ in range(pools):
            h, w = (h + 1) // 2, (w + 1) // 2
        shapes[name] = (h, w, c)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    maps = {n: torch.randn((BATCH, s[0], s[1], s[2]), generator=g, device=dev).clamp_(min=0).to(torch.bfloat16).contiguous()
            for n, s in shapes.items()}
    boxes = torch.from_numpy(np.stack([synth.make_boxes(NPROP, IM_H, IM_W, seed=100 + i, lo=24, hi=400) for i in range(BATCH)])).to(dev)
    counts = torch.full((BATCH,), NPROP, dtype=torch.int32, device=dev)
    hd, P = head, head.pooled
    mc = BATCH * NPROP
    pooled_src = [torch.empty((mc * P * P, c), dtype=torch.bfloat16, device=dev) for c in hd.src_channels] if args.unfused else None
    stages = ["rois", "pool3", "pool4", "pool5", "grn_concat", "conv_pool5", "fc6", "fc7", "cls_bbox", "select"]
    acc = {k: 0.0 for k in stages}
    m_rows = []

    def step(timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
        ev[0].record()
        det.prepare(boxes, counts)
        ev[1].record()
        off = 0
        for k, (name, sc, c) in enumerate(zip(hd.conv_names, hd.scales, hd.src_channels)):
            if args.unfused:
                ops.roi_pool(maps[name], det.rois, P, sc, layout="NHWC", n_rois=det.m_total, out=pooled_src[k].view(mc, P, P, c), staged=True)
            else:
                ops.roi_pool_grn(maps[name], det.rois, det.cat, off, P, sc, hd.grn_scale, n_rois=det.m_total)
            off += c
            ev[2 + k].record()
        if args.unfused:
            ops.grn_concat(pooled_src, hd.grn_scale, n_units=det.m_total, rows_per_unit=P * P, out=det.cat)
        ev[5].record()
        torch.mul(det.m_total, P * P, out=det.m_rows)
        ops.fc_forward(det.cat, hd.wc, hd.bc, _lib.ACT_RELU, m_live=det.m_rows, out=det.pool5.view(mc * P * P, hd.C))
        ev[6].record()
        ops.fc_forward(det.pool5, hd.w6, hd.b6, _lib.ACT_RELU, m_live=det.m_total, out=det.h6)
        ev[7].record()
        ops.fc_forward(det.h6, hd.w7, hd.b7, _lib.ACT_RELU, m_live=det.m_total, out=det.h7)
        ev[8].record()
        ops.fc_forward(det.h7, hd.wo, hd.bo, _lib.ACT_SOFTMAX_BBOX, det.C, m_live=det.m_total, out=det.out[:, :det.n_out])
        ev[9].record()
        det.select(**dset.slot(0, BATCH))
        ev[10].record()
        if timed:
            torch.cuda.synchronize()
            for k, name in enumerate(stages):
                acc[name] += ev[k].elapsed_time(ev[k + 1])
            m_rows.append(int(det.m_total.item()))

    for _ in range(max(args.warmup, 3)):
        step(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    for _ in range(args.steps):
        step(True)
    per = {k: round(v / args.steps, 4) for k, v in acc.items()}
    m = float(np.mean(m_rows))
    hbm_peak, tf_peak, which = B.peaks()
    ctot = sum(hd.src_channels)
    rate = {}
    for name, key in (("conv3_3", "pool3"), ("conv4_3", "pool4"), ("conv5_3", "pool5")):
        hh, ww, c = shapes[name]
        rate[key + "_gbs"] = (BATCH * hh * ww * c * 2 + m * (20 + 49 * c * 2)) / (per[key] * 1e-3) / 1e9
    if args.unfused:
        rate["grn_concat_gbs"] = (m * 49 * ctot * 2 * 2) / (per["grn_concat"] * 1e-3) / 1e9
    rate["conv_pool5_tflops"] = 2.0 * m * 49 * ctot * hd.C / (per["conv_pool5"] * 1e-3) / 1e12
    rate["fc6_tflops"] = 2.0 * m * hd.w6.shape[0] * hd.w6.shape[1] / (per["fc6"] * 1e-3) / 1e12
    finite = bool(torch.isfinite(det.out[:int(m), :det.C]).all().item())
    line = {"metric": "skip-layer Fast R-CNN head images/sec (detection step only, proposals given)", "value": BATCH / (ms / 1e3),
            "unit": "images/s", "ms_per_step": ms, "steps": args.steps, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "VGG16_skip detector head, voc_skip.yml (MAX_SIZE 800, DEDUP_BOXES 0.5), 64 synthetic 600x1000 "
                                   "images x 300 proposals, %d classes; %s" % (args.classes, "separate pools + GRN pass" if args.unfused else "GRN fused into the pools"),
                       "maps": {n: list(s) for n, s in shapes.items()}, "unique_rois_per_step": m, "scores_finite": finite},
            "per_stage_ms": per, "rates": {k: round(v, 1) for k, v in rate.items()},
            "peaks": {"hbm_gbs": hbm_peak, "bf16_tflops": tf_peak, "source": which}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
