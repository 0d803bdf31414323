#!/bin/bash
# N = 8 check of the peer-window route (weak + job 1024) and the same box's N = 1.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=8
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
timeout -k 10 240 $T bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>gpurun_out/r2ba_peer.err | tail -1 > gpurun_out/r2ba_weak_peer_n8.json
timeout -k 10 240 $T bench.py --gpus $N --job 1024 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2ba_peer.err | tail -1 > gpurun_out/r2ba_job_peer_n8.json
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 > gpurun_out/r2ba_weak_n1.json
python - <<'PY'
import json
for f in ("weak_peer_n8", "job_peer_n8", "weak_n1"):
    try:
        d = json.load(open("gpurun_out/r2ba_%s.json" % f)); print(f, round(d["value"]), round(d["ms_per_step"], 4), d["steps"], round(d["e2e"]["value"]), d.get("gather_ms"))
    except Exception as e:
        print(f, "missing", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r2ba_peer.err | tail -5
