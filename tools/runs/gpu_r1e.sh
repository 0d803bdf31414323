#!/bin/bash
# Round-1e GPU session: parity suites, smoke, bench, backbone bench, launch list + a full capture of the conv kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gpu tests"; timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/t_gpu.log; tail -5 gpurun_out/t_gpu.log; grep -E "^E  |Error" gpurun_out/t_gpu.log | head -40
echo "== smoke"; timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout -k 10 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-700
echo "== backbone"; timeout -k 10 600 python tools/backbone_bench.py 2>&1 | tail -3 | tee gpurun_out/backbone.log | cut -c1-3000
echo "== ncu backbone launch list + conv full capture"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_backbone.csv python tools/backbone_bench.py --steps 1 --warmup 1 --no-cudnn > gpurun_out/launches_backbone.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:fc_gemm_kernel -s 13 -c 13 -o gpurun_out/prof_conv -f python tools/backbone_bench.py --steps 1 --warmup 1 --no-cudnn > gpurun_out/prof_conv.log 2>&1
ls -la gpurun_out
