#!/bin/bash
# ROI pool, two-level map (pool_bins_pairs): parity tests + A/B microbench.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "roi_pool" 2>&1 | tail -5
for m in 22 422; do timeout 200 python tools/microbench.py --only roi_pool --pool-mode $m 2>&1 | grep nhwc | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('38x63 mode',d['pool_mode'],d['variant'],d['R'],round(d['ms_best'],4),round(d['frac_of_measured_hbm'],4))"; done
for m in 22 222 622; do timeout 200 python tools/microbench.py --only roi_pool --hw 30,50 --pool-mode $m 2>&1 | grep nhwc | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('30x50 mode',d['pool_mode'],d['variant'],d['R'],round(d['ms_best'],4),round(d['frac_of_measured_hbm'],4))"; done
