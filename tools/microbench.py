#!/usr/bin/env python
"""ROI-pool + NMS microbench sweep (BASELINE.json config #4): R, N in {2000, 8000, 20000}, IoU 0.3 / 0.7,
map 512x38x63.  Prints one JSON line per case: ROI-pool achieved HBM GB/s on the ALGORITHMIC bytes
(map read once + 20 B/ROI + pooled rows written once), NMS boxes/sec; CUDA events, best of 10 after 3
warm-ups, L2 flushed between iterations.  --cpu adds the oracle's single-core time for the same case."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aznet_b200 import _lib, ops, synth  # noqa: E402


def timeit(fn, flush, iters=10, warm=3):
    for _ in range(warm):
        fn()
    best, tot = 1e9, 0.0
    for _ in range(iters):
        flush.zero_()                                  # 256 MB write: evicts L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best, tot = min(best, ms), tot + ms
    return best, tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--sizes", default="2000,8000,20000")
    ap.add_argument("--hw", default="38,63", help="conv5_3 map size: 38,63 (default cfg, scale 1.0) or 30,50 (voc.yml, scale 0.8)")
    ap.add_argument("--only", default="", help="roi_pool | nms")
    ap.add_argument("--nms-phases", action="store_true", help="also time the NMS phases (azn_nms_tune): sort, sort+mask, sequential")
    ap.add_argument("--pool-mode", type=int, default=0, help="azn_roi_pool_tune: 0 auto, 1 direct, 2 staged")
    args = ap.parse_args()
    _lib.build()
    _lib.require_device()
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(os.path.dirname(_lib.HEADER), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(os.path.dirname(_lib.HEADER), "..", "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    C = 512
    H, W = [int(x) for x in args.hw.split(",")]
    feat = torch.from_numpy(synth.make_conv_maps(1, C, H, W, seed=7)).to(dev)
    nhwc32 = feat.permute(0, 2, 3, 1).contiguous()
    nhwc16 = nhwc32.to(torch.bfloat16)
    if args.cpu:
        from oracle import az_oracle as O
    _lib.lib().azn_roi_pool_tune(args.pool_mode)
    for R in [int(x) for x in args.sizes.split(",")] if args.only in ("", "roi_pool") else []:
        rois = torch.from_numpy(synth.make_rois(R, 600, 1000, seed=3)).to(dev)
        for name, f, layout, esz in (("nhwc_bf16", nhwc16, "NHWC", 2), ("nhwc_f32", nhwc32, "NHWC", 4), ("nchw_f32", feat, "NCHW", 4)):
            shape = (R, 7, 7, C) if layout == "NHWC" else (R, C, 7, 7)
            out = torch.empty(shape, dtype=f.dtype, device=dev)
            best, mean = timeit(lambda: ops.roi_pool(f, rois, layout=layout, out=out), flush)
            nbytes = C * H * W * esz + R * (20 + C * 49 * esz)
            line = {"bench": "roi_pool", "variant": name, "pool_mode": args.pool_mode, "map": [H, W], "R": R, "ms_best": best, "ms_mean": mean, "bytes": nbytes,
                    "gbs": nbytes / (best * 1e-3) / 1e9, "frac_of_measured_hbm": nbytes / (best * 1e-3) / 1e9 / peak}
            if args.cpu and name == "nchw_f32" and R <= 2000:
                t0 = time.perf_counter()
                O.roi_pool_fwd(feat.cpu().numpy(), rois.cpu().numpy())
                line["cpu_oracle_s"] = time.perf_counter() - t0
            print(json.dumps(line), flush=True)
    for N in [int(x) for x in args.sizes.split(",")] if args.only in ("", "nms") else []:
        d_host = synth.make_dets(N, seed=3)
        d = torch.from_numpy(d_host).to(dev)
        for th in (0.3, 0.7):
            res = {}

            def run():
                res["k"], res["c"] = ops.nms(d, th)
            best, mean = timeit(run, flush)
            line = {"bench": "nms", "N": N, "thresh": th, "ms_best": best, "ms_mean": mean, "kept": int(res["c"].item()),
                    "boxes_per_s": N / (best * 1e-3), "mask_bytes": 16 * N * ((N + 63) // 64) + 28 * N}
            if args.nms_phases:
                ph = {}
                for mode, name in ((3, "sort_ms"), (8 + 3, "sort_allpairs_ms"), (2, "sort_mask_ms"), (1, "sequential_ms"), (8, "total_allpairs_sort_ms")):
                    _lib.lib().azn_nms_tune(mode)
                    ph[name] = round(timeit(run, flush)[0], 4)
                _lib.lib().azn_nms_tune(0)
                line["phases"] = ph
            if args.cpu and N <= 8000:
                t0 = time.perf_counter()
                k = O.nms(d_host, th)
                line["cpu_oracle_s"] = time.perf_counter() - t0
                line["cpu_boxes_per_s"] = N / line["cpu_oracle_s"]
                line["match"] = k == res["k"][:len(k)].cpu().tolist()
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
