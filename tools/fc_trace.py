#!/usr/bin/env python
"""In-kernel timeline of azn_fc_forward (azn_fc_trace hook): per layer shape and live row count, the time
each phase of the persistent GEMM takes -- setup, first operands landed, main loop, epilogue / fix-up,
teardown -- as min / mean / max over the CTAs that had work, from %globaltimer stamps (ns)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aznet_b200 import _lib as L, ops  # noqa: E402


def main():
    L.build()
    L.require_device()
    dev = torch.device("cuda:0")
    layers = [("int6", 4096, 25088, L.ACT_RELU, 0, torch.bfloat16), ("int7", 1280, 4096, L.ACT_RELU, 0, torch.bfloat16),
              ("heads", 56, 1280, L.ACT_AZ_HEAD, 11, torch.float32)]
    Ms = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "64,512,1500").split(",")]
    tiles = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
    cap = max(Ms)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    tr = torch.zeros((sms, 8), dtype=torch.int64, device=dev)
    for name, N, K, act, aux, odt in layers:
        A = (torch.randn((cap, K), device=dev) * 0.1).to(torch.bfloat16)
        W = (torch.randn((N, K), device=dev) * 0.01).to(torch.bfloat16)
        b = torch.zeros(N, device=dev)
        out = torch.empty((cap, (N + 7) // 8 * 8), dtype=odt, device=dev)[:, :N] if odt == torch.float32 else torch.empty((cap, N), dtype=odt, device=dev)
        for M in Ms:
            ml = torch.tensor([M], dtype=torch.int32, device=dev)
            for tile in tiles:
                L.lib().azn_fc_tune(0, -1, tile)
                for _ in range(3):
                    ops.fc_forward(A, W, b, act, aux, m_live=ml, out=out)
                torch.cuda.synchronize()
                L.lib().azn_fc_trace(C.c_void_p(tr.data_ptr()))
                tr.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.fc_forward(A, W, b, act, aux, m_live=ml, out=out)
                e1.record()
                torch.cuda.synchronize()
                L.lib().azn_fc_trace(None)
                t = tr.cpu().numpy().astype("int64")
                t0 = t[:, 0].min()
                busy = t[:, 7] > 0
                row = {"layer": name, "M": M, "tile": tile, "event_us": round(e0.elapsed_time(e1) * 1e3, 1), "ctas_busy": int(busy.sum()),
                       "units_max": int(t[:, 7].max())}

                def st(x):
                    x = x[busy] if busy.any() else x
                    return [round(float(x.min()) / 1e3, 1), round(float(x.mean()) / 1e3, 1), round(float(x.max()) / 1e3, 1)]
                row["start_skew_us"] = st(t[:, 0] - t0)
                row["setup_us"] = st(t[:, 1] - t[:, 0])
                row["first_full_us"] = st(t[:, 2] - t[:, 1])
                row["mainloop_us"] = st(t[:, 3] - t[:, 2])
                row["acc_ready_after_start_us"] = st(t[:, 4] - t[:, 0])
                row["epilogue_after_mma_us"] = st(t[:, 5] - t[:, 3])
                row["teardown_us"] = st(t[:, 6] - t[:, 5])
                row["kernel_span_us"] = round(float(t[:, 6].max() - t0) / 1e3, 1)
                print(json.dumps(row), flush=True)
    L.lib().azn_fc_tune(0, -1, 0)


if __name__ == "__main__":
    main()
