#!/bin/bash
# ncu --set full of the default staged ROI-pool kernel after the fix (R = 20 000 bf16), 38x63 and 30x50 maps.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for hw in 38,63 30,50; do
  tag=${hw/,/x}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_pool_keys_kernel -s 3 -c 1 -f -o gpurun_out/r2br_pool_$tag python tools/microbench.py --only roi_pool --hw $hw --sizes 20000 > gpurun_out/r2br_ncu_$tag.log 2>&1
done
ls -la gpurun_out/r2br*
