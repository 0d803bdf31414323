#!/bin/bash
# Peer-window gather (round 2, last session): test on one GPU, then A/B peer vs nccl at N = 2 under torchrun, weak + job 1024.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py -x -q 2>&1 | tail -5
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
for G in peer nccl; do
  echo "== weak N=$N gather=$G"; timeout -k 10 300 $T bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extra --gather $G 2>gpurun_out/r2at_${G}_n$N.err | tail -1 | tee gpurun_out/r2at_weak_${G}_n$N.json | cut -c1-150
  echo "== job N=$N gather=$G"; timeout -k 10 300 $T bench.py --gpus $N --job 1024 --warmup 3 --no-cpu-baseline --no-extra --gather $G 2>>gpurun_out/r2at_${G}_n$N.err | tail -1 | tee gpurun_out/r2at_job_${G}_n$N.json | cut -c1-150
done
echo "== weak N=1"; timeout -k 10 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>gpurun_out/r2at_n1.err | tail -1 | tee gpurun_out/r2at_weak_n1.json | cut -c1-150
python - <<'PY'
import json
for f in ("weak_peer_n2", "weak_nccl_n2", "job_peer_n2", "job_nccl_n2", "weak_n1"):
    try:
        d = json.load(open("gpurun_out/r2at_%s.json" % f))
        print(f, round(d["value"]), round(d["ms_per_step"], 4), d["steps"], "e2e", round(d["e2e"]["value"]), d.get("gather_ms"))
    except Exception as e:
        print(f, "missing", e)
PY
tail -5 gpurun_out/r2at_peer_n$N.err
