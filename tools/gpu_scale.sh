#!/bin/bash
# Weak-scaling run on an 8-GPU B200 box (`gpurun --gpus 8`): the bench under torchrun exactly as the driver launches it,
# at N = 8, 4, 2, 1 (64 images per GPU per step).  Logs -> gpurun_out/scale_n*.log.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpus_scale.txt 2>&1
for N in ${@:-8 4 2}; do
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N))"
  echo "== bench N=$N"; timeout -k 10 300 $T bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/scale_n$N.log | cut -c1-330
done
echo "== bench N=1"; timeout -k 10 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/scale_n1.log | cut -c1-330
