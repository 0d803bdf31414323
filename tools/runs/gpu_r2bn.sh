#!/bin/bash
# Re-entry check at HEAD on 2 GPUs: both arms under torchrun exactly as the driver launches them (default flags).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
t0=$(date +%s)
timeout -k 10 300 $T bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2bn_ref_n2.json; cut -c1-200 gpurun_out/r2bn_ref_n2.json
t1=$(date +%s); echo "reference arm: $((t1-t0)) s"
timeout -k 10 400 $T bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r2bn_n2.err | tail -1 > gpurun_out/r2bn_bench_n2.json
t2=$(date +%s); echo "bench N=2: $((t2-t1)) s"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2bn_bench_n2.json"))
print(d["value"], d["ms_per_step"], d["n_gpus"], d["gpu_launches"], d["clocks"], d.get("gather_ms"))
print(d["e2e"]["value"], d["e2e"].get("route_timing",{}).get("chosen"), d["roofline"]["frac"])
PY
tail -3 gpurun_out/r2bn_n2.err
