#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command (host-launched so that every kernel is a separate
# launch) + full captures of the top kernels.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_bench.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:fc_gemm_kernel -s 54 -c 6 -o gpurun_out/prof_gemm -f $B > gpurun_out/prof_gemm.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:roi_pool_nhwc -s 18 -c 2 -o gpurun_out/prof_pool -f $B > gpurun_out/prof_pool.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"search_level|select_kernel" -s 8 -c 3 -o gpurun_out/prof_search -f $B > gpurun_out/prof_search.log 2>&1
ls -la gpurun_out
# the microbench kernels (BASELINE config #4): staged ROI pool and NMS at the largest size
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"roi_pool_keys|nms_" -s 3 -c 5 -o gpurun_out/prof_micro -f python tools/microbench.py --sizes 20000 > gpurun_out/prof_micro.log 2>&1
ls -la gpurun_out
