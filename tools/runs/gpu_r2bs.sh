#!/bin/bash
# A/B: the search levels' ROI pool through the staged kernel (per-image buckets) instead of the direct one.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lv in "" 2 2,3 2,3,4,5 ""; do
  timeout 200 python bench.py --steps 12 --warmup 3 --no-extra --no-cpu-baseline --pool-staged-levels "$lv" 2>gpurun_out/r2bs.err | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
pl = [(x['level'], x['ms']) for x in d['roofline']['per_level'] if x['stage'] == 'roi_pool']
print('staged levels [$lv]', round(d['value']), round(d['ms_per_step'], 4), 'pool ms per level', pl, 'e2e', round(d['e2e']['value']))
" | tee -a gpurun_out/r2bs_ab.txt
  tail -1 gpurun_out/r2bs.err | cut -c1-200
done
