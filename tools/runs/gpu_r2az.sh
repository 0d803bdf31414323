#!/bin/bash
# ROI pool, geometry pre-pass (roi_geom_kernel): parity tests + A/B microbench (8xx = without the pre-pass).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -x -q -k "roi_pool" 2>&1 | tail -5
for hw in 38,63 30,50; do for m in 22 822 22 822; do timeout 200 python tools/microbench.py --only roi_pool --hw $hw --pool-mode $m 2>&1 | grep nhwc | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$hw mode',d['pool_mode'],d['variant'],d['R'],round(d['ms_best'],4),round(d['frac_of_measured_hbm'],4))"; done; done
