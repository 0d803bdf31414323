#!/usr/bin/env python
"""SURVEY 8f-1 measurement: the VGG16 conv1_1 .. conv5_3 backbone on hand-written sm_100a kernels
(azn_image_blob, azn_conv3x3_forward, azn_maxpool2x2_forward), batch of synthetic 600x1000 uint8 images at the
PASCAL scale (MAX_SIZE 800 -> 480x800 network input -> conv5_3 512x30x50).

Per layer: CUDA-event time, algorithmic FLOPs (real channel counts, unpadded pixels) and TFLOP/s against the
measured sustained bf16 peak; whole backbone: images/s; and, as a LIBRARY comparison only, the same stack through
PyTorch/cuDNN bf16 channels-last.  Prints one JSON line.

    python tools/backbone_bench.py [--batch 16] [--steps 10] [--warmup 3] [--no-cudnn]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

IM_H, IM_W = 600, 1000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--max-size", type=int, default=800)
    ap.add_argument("--no-cudnn", action="store_true")
    args = ap.parse_args()
    import torch
    from aznet_b200 import _lib, backbone, engine, ops, synth
    import bench as B

    _lib.build()
    _lib.require_device()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    _, peak_tf, _ = B.peaks()
    scale = engine.im_scale_for(IM_H, IM_W, max_size=args.max_size)
    hs, ws = ops.blob_size(IM_H, IM_W, scale)
    w = backbone.make_vgg16_weights(seed=5)
    net = backbone.VGG16Native(w, dev)
    # two distinct batches so that no step finds its input in L2 (the maps themselves are far larger than L2)
    ims = [torch.from_numpy(np.stack(synth.make_images(args.batch, IM_H, IM_W, seed=1000 + 100 * s))).to(dev) for s in range(2)]
    direct = net.first_patches and net.first_direct          # conv1_1 in one mma.sync kernel, no patch matrix
    names = ["blob"] + (["patches"] if net.first_patches and not direct else [])
    for s, (_, n) in enumerate(backbone.VGG16_CFG, 1):
        for i in range(1, n + 1):
            names.append("conv%d_%d" % (s, i))
            if i == n and s < 5:
                names.append("pool%d" % s)
    acc = {k: 0.0 for k in names}

    def step(i, timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        ev[0].record()
        x = ops.image_blob(ims[i % 2], scale, net.pixel_means, net.cpad_in)
        ev[1].record()
        k, last = 1, len(net.layers) - 1
        for li, (wt, b, pool) in enumerate(net.layers):
            if li == 0 and direct:
                x = ops.conv_direct(x, net.in_channels, wt, b, relu=True)
            elif li == 0 and net.first_patches:
                x = ops.patches3x3(x, net.in_channels, 64)
                k += 1
                ev[k].record()
                x = ops.conv_patches(x, wt, b, relu=True, unpadded=(li == last))
            else:
                x = ops.conv3x3(x, wt, b, relu=True, unpadded=(li == last))
            k += 1
            ev[k].record()
            if pool:
                x = ops.maxpool2x2(x)
                k += 1
                ev[k].record()
        if timed:
            torch.cuda.synchronize()
            for j, nm in enumerate(names):
                acc[nm] += ev[j].elapsed_time(ev[j + 1])
        return x

    for i in range(args.warmup):
        out = step(i, False)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        out = net.from_images(ims[i % 2], scale)
    t1.record()
    torch.cuda.synchronize()
    ms_step = t0.elapsed_time(t1) / args.steps
    for i in range(args.steps):
        step(i, True)
    # per-layer flops (real channels, unpadded pixels)
    layers, h, wd, li = [], hs, ws, 0
    total_flops = 0.0
    for nm in names:
        ms = acc[nm] / args.steps
        if nm.startswith("conv"):
            ci, co, pool = net.dims[li]
            li += 1
            fl = 2.0 * 9 * ci * co * h * wd * args.batch
            total_flops += fl
            layers.append({"layer": nm, "ms": round(ms, 4), "cin": ci, "cout": co, "hw": [h, wd], "tflops": round(fl / ms / 1e9, 1),
                           "frac": round(fl / ms / 1e9 / peak_tf, 3)})
        elif nm.startswith("pool"):
            byts = args.batch * h * wd * net.dims[li - 1][1] * 2 * 1.25     # read once + quarter-size write
            layers.append({"layer": nm, "ms": round(ms, 4), "gbs": round(byts / ms / 1e6, 1)})
            h, wd = (h + 1) // 2, (wd + 1) // 2
        elif nm == "patches":
            byts = args.batch * (hs + 2) * (ws + 2) * (net.cpad_in + 64) * 2
            layers.append({"layer": nm, "ms": round(ms, 4), "gbs": round(byts / ms / 1e6, 1)})
        else:
            byts = args.batch * (IM_H * IM_W * 3 + (hs + 2) * (ws + 2) * net.cpad_in * 2)
            layers.append({"layer": nm, "ms": round(ms, 4), "gbs": round(byts / ms / 1e6, 1)})
    line = {"metric": "VGG16 conv5_3 backbone images/sec", "value": args.batch / ms_step * 1e3, "unit": "images/s",
            "ms_per_step": ms_step, "batch": args.batch, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "VGG16 conv1_1..conv5_3, %d synthetic %dx%d uint8 images -> %dx%d input -> conv5_3 %s"
                                   % (args.batch, IM_H, IM_W, hs, ws, list(out.shape))},
            "gflop_per_image": total_flops / args.batch / 1e9,
            "tflops": total_flops / ms_step / 1e9, "frac_of_sustained_bf16": total_flops / ms_step / 1e9 / peak_tf,
            "peak_tflops": peak_tf, "gpu_launches_per_step": net.launches_per_call, "layers": layers}
    if not args.no_cudnn:
        ref = backbone.VGG16Torch(w, dev)
        data = torch.randn((args.batch, 3, hs, ws), device=dev)
        for _ in range(args.warmup):
            ref(data)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(args.steps):
            ref(data)
        t1.record()
        torch.cuda.synchronize()
        ms_ref = t0.elapsed_time(t1) / args.steps
        line["cudnn_library_reference"] = {"ms_per_step": ms_ref, "images_per_s": args.batch / ms_ref * 1e3,
                                           "tflops": total_flops / ms_ref / 1e9}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
