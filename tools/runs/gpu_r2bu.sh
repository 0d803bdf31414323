#!/bin/bash
# Search engine: ROI pool with one CTA per ROI from level 4 (the new default) against never (--pool-per-roi-from 0) and
# from level 3 / 5; full GPU suite first.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
for from in 0 4 5 0 4; do
  timeout 200 python bench.py --steps 12 --warmup 3 --no-extra --no-cpu-baseline --pool-per-roi-from $from 2>gpurun_out/r2bu.err | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
pl = [(x['level'], x['ms']) for x in d['roofline']['per_level'] if x['stage'] == 'roi_pool']
print('per-ROI pool from level $from:', round(d['value']), round(d['ms_per_step'], 4), 'pool ms per level', pl)
" | tee -a gpurun_out/r2bu_ab.txt
  tail -1 gpurun_out/r2bu.err | cut -c1-200
done
