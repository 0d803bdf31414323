cd $GRAFT_REPO_ROOT
AZN_NMS_NO_PARTITION=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_ -c 40 -o gpurun_out/r2ak_nms -f python tools/microbench.py --only nms --sizes 20000 > gpurun_out/r2ak.log 2>&1
ls -la gpurun_out/r2ak*; tail -2 gpurun_out/r2ak.log
timeout 300 python tools/microbench.py --sizes 2000,8000,20000 --nms-phases > gpurun_out/r2ak_microbench_3863.jsonl 2>&1
timeout 300 python tools/microbench.py --sizes 2000,8000,20000 --hw 30,50 --only roi_pool > gpurun_out/r2ak_microbench_3050.jsonl 2>&1
tail -2 gpurun_out/r2ak_microbench_3863.jsonl
