#!/bin/bash
# Runs on the B200 box through gpurun: GPU parity suites, smoke, a short bench. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernels (non-GEMM)"; timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not fc_forward" 2>&1 | tail -25 | tee gpurun_out/t_kernels.log
echo "== fc_forward"; timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fc_forward" 2>&1 | tail -40 | tee gpurun_out/t_fc.log
echo "== search"; timeout -k 10 600 python -m pytest tests/test_gpu_search.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/t_search.log
echo "== smoke"; timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "== bench"; timeout -k 10 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
