#!/usr/bin/env python
"""Tuning sweep for azn_fc_forward: layer shapes of the AZ head x live row counts x (tile width, split factor,
fix-up mode).  CUDA events, best of 8 after 2 warm-ups, weights larger than L2 are naturally cold."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aznet_b200 import _lib as L, ops  # noqa: E402


def run(A, W, b, act, aux, out, ml):
    best = 1e9
    for it in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.fc_forward(A, W, b, act, aux, m_live=ml, out=out)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            best = min(best, e0.elapsed_time(e1))
    return best


def main():
    L.build()
    L.require_device()
    dev = torch.device("cuda:0")
    layers = [("int6", 4096, 25088, L.ACT_RELU, 0, torch.bfloat16), ("int7", 1280, 4096, L.ACT_RELU, 0, torch.bfloat16),
              ("heads", 56, 1280, L.ACT_AZ_HEAD, 11, torch.float32)]
    Ms = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "64,512,880,1500,4000").split(",")]
    cap = max(Ms)
    for name, N, K, act, aux, odt in layers:
        A = (torch.randn((cap, K), device=dev) * 0.1).to(torch.bfloat16)
        W = (torch.randn((N, K), device=dev) * 0.01).to(torch.bfloat16)
        b = torch.zeros(N, device=dev)
        out = torch.empty((cap, (N + 7) // 8 * 8), dtype=odt, device=dev)[:, :N] if odt == torch.float32 else torch.empty((cap, N), dtype=odt, device=dev)
        for M in Ms:
            ml = torch.tensor([M], dtype=torch.int32, device=dev)
            res = {}
            bns = [2256, 1256, 2128, 1128] if N >= 1024 else [64, 1128, 2128]      # tile code: 1000 * row halves + width
            for bn in bns:
                for parts, mode in [(0, -1), (1, 0), (2, 0), (3, 0), (4, 0), (2, 1), (4, 1), (8, 1)]:
                    L.lib().azn_fc_tune(parts, mode, bn)
                    res["bn%d_p%d_m%d" % (bn, parts, mode)] = round(run(A, W, b, act, aux, out, ml) * 1e3, 1)
            L.lib().azn_fc_tune(0, -1, 0)
            res["auto"] = round(run(A, W, b, act, aux, out, ml) * 1e3, 1)
            print(json.dumps({"layer": name, "M": M, "us": res}), flush=True)


if __name__ == "__main__":
    main()
