#!/usr/bin/env python
"""Where does the fixed cost of a small azn_fc_forward go?  Times (CUDA events, min of 20) the call with
m_live = 0 (prologue + teardown only), 1 k-block, and growing K for one 128-row tile, plus back-to-back calls."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aznet_b200 import _lib as L, ops


def t(fn, n=20, reps=1):
    best = 1e9
    for it in range(n + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            best = min(best, e0.elapsed_time(e1) / reps)
    return round(best * 1e3, 2)


def main():
    L.build(); L.require_device()
    dev = torch.device("cuda:0")
    x = torch.zeros(1024, device=dev)
    print(json.dumps({"empty_torch_kernel_us": t(lambda: x.add_(1.0)), "x10": t(lambda: x.add_(1.0), reps=10)}))
    for N in (56, 1280, 4096):
        for K in (64, 256, 1280, 4096):
            A = (torch.randn((256, K), device=dev) * 0.1).to(torch.bfloat16)
            W = (torch.randn((N, K), device=dev) * 0.01).to(torch.bfloat16)
            b = torch.zeros(N, device=dev)
            out = torch.empty((256, N), dtype=torch.bfloat16, device=dev)
            row = {"N": N, "K": K}
            for M in (0, 64, 256):
                ml = torch.tensor([M], dtype=torch.int32, device=dev)
                L.lib().azn_fc_tune(1, 0, 0)
                row["M%d_1x" % M] = t(lambda: ops.fc_forward(A, W, b, L.ACT_RELU, 0, m_live=ml, out=out))
                row["M%d_10x" % M] = t(lambda: ops.fc_forward(A, W, b, L.ACT_RELU, 0, m_live=ml, out=out), reps=10)
            # static M (no finish launch at all when tiles % grid == 0 is impossible here; m_live=None still launches finish)
            L.lib().azn_fc_tune(0, -1, 0)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
