#!/bin/bash
# Round-1g GPU session (final evidence of the round): all parity suites, smoke, bench, microbench (both map sizes),
# config #3 detbench, backbone bench, skip bench, and the ncu passes: launch list of the bench command + full captures
# of the GEMM / ROI-pool / search kernels (tools/gpu_profile.sh) and of the pipelined NMS.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gpu tests"; timeout -k 10 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/t_gpu.log; tail -5 gpurun_out/t_gpu.log; grep -E "^E  |Error|FAILED" gpurun_out/t_gpu.log | head -40
echo "== smoke"; timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench"; timeout -k 10 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-400
echo "== bench reference arm"; timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.log | cut -c1-400
echo "== microbench"; (timeout -k 10 600 python tools/microbench.py --nms-phases; timeout -k 10 300 python tools/microbench.py --only roi_pool --hw 30,50) 2>&1 | tee gpurun_out/microbench.jsonl | cut -c1-160
echo "== detbench"; timeout -k 10 600 python tools/detbench.py 2>&1 | tail -1 | tee gpurun_out/detbench.json | cut -c1-400
echo "== backbone"; timeout -k 10 600 python tools/backbone_bench.py 2>&1 | tail -1 | tee gpurun_out/backbone.json | cut -c1-300
echo "== skipbench"; timeout -k 10 600 python tools/skipbench.py 2>&1 | tail -1 | tee gpurun_out/skipbench.json | cut -c1-300
echo "== ncu"; bash tools/gpu_profile.sh > gpurun_out/profile.log 2>&1
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:"nms_" -s 60 -c 30 -o gpurun_out/prof_nms -f python tools/microbench.py --sizes 20000 --only nms > gpurun_out/prof_nms.log 2>&1
ls -la gpurun_out | head -40
# summaries on the box (the .ncu-rep files together exceed what gpurun merges back)
for r in gemm pool search micro nms; do
  [ -f gpurun_out/prof_$r.ncu-rep ] && python tools/ncu_summary.py gpurun_out/prof_$r.ncu-rep gpurun_out/ncu_$r.csv && rm -f gpurun_out/prof_$r.ncu-rep
done
ls -la gpurun_out | head -40
