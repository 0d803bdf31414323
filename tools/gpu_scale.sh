#!/bin/bash
# Scaling runs on an 8-GPU B200 box (`gpurun --gpus 8`): the bench under torchrun exactly as the driver launches it.
#   weak   64 images per GPU per step, 10 steps                          -> gpurun_out/r2_scale_weak_n*.json
#   strong BASELINE config #5 as written: a job of 1024 images = 16 batches of 64 sharded over the ranks (--job 1024)
#                                                                         -> gpurun_out/r2_scale_job1024_n*.json
# plus the PCIe / NVLink topology (`nvidia-smi topo -m`) that bounds the host-fed e2e figure.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,pci.bus_id --format=csv > gpurun_out/r2_gpus_scale.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r2_gpus_scale.txt 2>&1
lscpu | grep -i "numa\|model name\|^cpu(s)" >> gpurun_out/r2_gpus_scale.txt 2>&1
for N in ${@:-8 4 2}; do
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N))"
  echo "== weak N=$N"; timeout -k 10 300 $T bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>gpurun_out/r2_scale_n$N.err | tail -1 | tee gpurun_out/r2_scale_weak_n$N.json | cut -c1-200
  echo "== job 1024 N=$N"; timeout -k 10 300 $T bench.py --gpus $N --job 1024 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2_scale_n$N.err | tail -1 | tee gpurun_out/r2_scale_job1024_n$N.json | cut -c1-200
done
echo "== weak N=1"; timeout -k 10 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>gpurun_out/r2_scale_n1.err | tail -1 | tee gpurun_out/r2_scale_weak_n1.json | cut -c1-200
echo "== job 1024 N=1"; timeout -k 10 300 python bench.py --gpus 1 --job 1024 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2_scale_n1.err | tail -1 | tee gpurun_out/r2_scale_job1024_n1.json | cut -c1-200
python - <<'PY'
import json, glob
for kind in ("weak", "job1024"):
    for n in (1, 2, 4, 8):
        try:
            d = json.load(open("gpurun_out/r2_scale_%s_n%d.json" % (kind, n)))
            print(kind, n, round(d["value"]), round(d["ms_per_step"], 4), d["steps"], "e2e", round(d["e2e"]["value"]), d.get("gather_ms"))
        except Exception as e:
            print(kind, n, "missing", e)
PY
