set -x
cd $GRAFT_REPO_ROOT
python - <<'PY' > gpurun_out/r2e_int7_sweep.jsonl 2>&1
import json, sys, torch
sys.path.insert(0, ".")
from aznet_b200 import _lib as L, ops
sys.path.insert(0, "tools")
import fc_sweep
L.build(); L.require_device()
dev = torch.device("cuda:0")
N, K = 1280, 4096
cap = 1536
A = (torch.randn((cap, K), device=dev) * 0.1).to(torch.bfloat16)
W = (torch.randn((N, K), device=dev) * 0.01).to(torch.bfloat16)
b = torch.zeros(N, device=dev)
out = torch.empty((cap, N), dtype=torch.bfloat16, device=dev)
for M in (576, 609, 878, 1494):
    ml = torch.tensor([M], dtype=torch.int32, device=dev)
    res = {}
    for bn in (1128, 2128, 1256, 2256, 1064):
        for parts, mode in [(0, -1), (1, 0), (2, 0), (4, 0), (2, 1), (4, 1), (8, 1)]:
            L.lib().azn_fc_tune(parts, mode, bn)
            res["bn%d_p%d_m%d" % (bn, parts, mode)] = round(fc_sweep.run(A, W, b, L.ACT_RELU, 0, out, ml) * 1e3, 1)
    L.lib().azn_fc_tune(0, -1, 0)
    res["auto"] = round(fc_sweep.run(A, W, b, L.ACT_RELU, 0, out, ml) * 1e3, 1)
    print(json.dumps({"layer": "int7", "M": M, "us": res}), flush=True)
PY
cat gpurun_out/r2e_int7_sweep.jsonl | cut -c1-1500
