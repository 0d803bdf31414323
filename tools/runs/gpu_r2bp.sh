#!/bin/bash
# ROI pool after moving the pair-table and geometry-pre-pass code out of the default instantiation (EXT = 3 / 4):
# parity tests of every variant, then the microbench A/B default vs 922 (geometry pre-pass) vs 422 (pair table).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_skip.py -m gpu -q -k "roi_pool or pool" 2>&1 | tail -3
for mode in 0 922 422 0; do
  for hw in 38,63 30,50; do
    timeout 200 python tools/microbench.py --only roi_pool --hw $hw --pool-mode $mode 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d['variant'] != 'nchw_f32': print('mode$mode', d['variant'], d['map'], d['R'], round(d['ms_best'], 4), round(d['ms_mean'], 4), round(d['frac_of_measured_hbm'], 3))
" | tee -a gpurun_out/r2bp_pool_ab.txt
  done
done
