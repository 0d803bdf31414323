#!/bin/bash
# Runs on an N-GPU B200 box through `gpurun --gpus N`: the bench under torchrun exactly as the driver launches it
# (our arm and the reference arm), plus the 1-GPU line for the scaling ratio.  Logs -> gpurun_out/.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpus_multi.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== reference arm under torchrun"; timeout -k 10 600 $T bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref_n$N.log | cut -c1-400
echo "== bench N=$N"; timeout -k 10 900 $T bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_n$N.log | cut -c1-700
echo "== bench N=1"; timeout -k 10 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n1.log | cut -c1-300
