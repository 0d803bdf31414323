set -x
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_pool_keys -s 3 -c 1 -o gpurun_out/r2w_pool_grouped -f python tools/microbench.py --only roi_pool --pool-mode 30 --sizes 20000 --hw 38,63 > gpurun_out/r2w.log 2>&1
ls -la gpurun_out/r2w*
