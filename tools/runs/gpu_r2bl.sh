#!/bin/bash
# Final launch lists of round 2: the bench command (host-launched, --no-extra so that only the headline path runs) and
# one backbone pass with the RU convolution kernels.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2bl_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extra > gpurun_out/r2bl_bench.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2bl_launches_backbone.csv python tools/backbone_bench.py --no-cudnn --batch 16 > gpurun_out/r2bl_backbone.log 2>&1
wc -l gpurun_out/r2bl_launches_bench.csv gpurun_out/r2bl_launches_backbone.csv
