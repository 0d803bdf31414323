#!/bin/bash
# ncu --set full of the convolution launches of one backbone pass (RU kernels), after warm-up
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fc_gemm_kernel -s 26 -c 13 -f -o gpurun_out/r2bh_conv python tools/backbone_bench.py --no-cudnn --batch 16 > gpurun_out/r2bh_ncu.log 2>&1
ls -la gpurun_out/r2bh*
