set -x
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_mask2_kernel -s 2 -c 1 -o gpurun_out/r2ad_nms_mask2 -f python tools/microbench.py --only nms --sizes 20000 > gpurun_out/r2ad.log 2>&1
ls -la gpurun_out/r2ad*
tail -3 gpurun_out/r2ad.log
