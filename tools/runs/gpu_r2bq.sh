#!/bin/bash
# ROI pool A/B: pool_bins_fixed with its seven-bin loop unrolled (default build) vs not unrolled (-DAZN_POOL_PW_UNROLL=1,
# aznet_b200/build/ab/lib_u1.so: the default kernel shrinks from 8880 to 3624 SASS instructions).  Same box, two rounds.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp aznet_b200/libaznet_b200.so /tmp/lib_head.so
for v in head u1 head u1; do
  if [ $v = head ]; then cp /tmp/lib_head.so aznet_b200/libaznet_b200.so; else cp aznet_b200/build/ab/lib_$v.so aznet_b200/libaznet_b200.so; fi
  touch aznet_b200/libaznet_b200.so
  for hw in 38,63 30,50; do
    timeout 200 python tools/microbench.py --only roi_pool --hw $hw 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d['variant'] != 'nchw_f32': print('$v', d['variant'], d['map'], d['R'], round(d['ms_best'], 4), round(d['ms_mean'], 4), round(d['frac_of_measured_hbm'], 3))
" | tee -a gpurun_out/r2bq_pool_ab.txt
  done
done
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -q -k "roi_pool" 2>&1 | tail -2
cp /tmp/lib_head.so aznet_b200/libaznet_b200.so
