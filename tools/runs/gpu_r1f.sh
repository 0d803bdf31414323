#!/bin/bash
# Round-1f GPU session: parity suites (incl. the tune row), smoke, bench, ncu captures of the CURRENT NMS kernels
# (microbench, N = 20 000) and of the detection step (config #3): launch list + full captures of the detect_* kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gpu tests"; timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/t_gpu.log; tail -5 gpurun_out/t_gpu.log; grep -E "^E  |Error" gpurun_out/t_gpu.log | head -40
echo "== smoke"; timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench"; timeout -k 10 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-700
echo "== ncu nms"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:"nms_" -s 60 -c 30 -o gpurun_out/prof_nms -f python tools/microbench.py --sizes 20000 --only nms > gpurun_out/prof_nms.log 2>&1
echo "== ncu detection"
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_det.csv python tools/detbench.py --steps 1 --warmup 1 --cpu-images 0 > gpurun_out/launches_det.log 2>&1
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:"detect_|nms_batched|nms_seg" -c 12 -o gpurun_out/prof_det -f python tools/detbench.py --steps 1 --warmup 1 --cpu-images 0 > gpurun_out/prof_det.log 2>&1
ls -la gpurun_out
