#!/bin/bash
cd $GRAFT_REPO_ROOT
for e in 0 1; do
  if [ $e = 1 ]; then export AZN_POOL_BIGSMEM=1; fi
  timeout 200 python tools/microbench.py --only roi_pool --pool-mode 422 --sizes 20000 2>&1 | grep nhwc | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('bigsmem $e mode',d['pool_mode'],d['variant'],d['R'],round(d['ms_best'],4),round(d['frac_of_measured_hbm'],4))"
done
