#!/bin/bash
# Runs on the B200 box through gpurun: GPU parity suites, smoke, a short bench, the microbench. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== fc tests first (a scheduling bug in the persistent GEMM would hang: short leash)"
timeout -k 10 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k fc_forward 2>&1 | tail -30 | tee gpurun_out/t_fc.log
if ! grep -q "passed" gpurun_out/t_fc.log || grep -q "failed" gpurun_out/t_fc.log; then echo "fc tests did not pass; stopping"; exit 1; fi
echo "== gpu tests"; timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/t_gpu.log; tail -5 gpurun_out/t_gpu.log; grep -E "^E  |Error" gpurun_out/t_gpu.log | head -40
echo "== smoke"; timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "== bench"; timeout -k 10 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log | cut -c1-600
echo "== bench eager"; timeout -k 10 900 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_eager.log | cut -c1-300
if [ "$1" == "micro" ]; then echo "== microbench"; timeout -k 10 600 python tools/microbench.py 2>&1 | tee gpurun_out/microbench.jsonl | cut -c1-250; fi
if [ "$1" == "prof" ] || [ "$2" == "prof" ]; then bash tools/gpu_profile.sh; fi
