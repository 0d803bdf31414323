#!/bin/bash
# ROI pool: ncu --set full of the two-level-map variant (22) and of the slice-only variant (422), R = 20 000 bf16 38x63.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in 22 422; do
ncu --set full --clock-control none --import-source on -k regex:roi_pool_keys_kernel -s 3 -c 1 -f -o gpurun_out/r2av_pool_$m python tools/microbench.py --only roi_pool --pool-mode $m --sizes 20000 > gpurun_out/r2av_ncu_$m.log 2>&1
done
ls -la gpurun_out/r2av*
