#!/bin/bash
# Direct ROI-pool kernel with one CTA per ROI (azn_roi_pool_tune(5xx)): parity of the variant, then the bench A/B.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -q -k "roi_pool" 2>&1 | tail -3
for tune in -1 500 -1 500; do
  timeout 200 python bench.py --steps 12 --warmup 3 --no-extra --no-cpu-baseline --pool-tune $tune 2>gpurun_out/r2bt.err | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
pl = [(x['level'], x['ms']) for x in d['roofline']['per_level'] if x['stage'] == 'roi_pool']
print('pool tune $tune', round(d['value']), round(d['ms_per_step'], 4), 'pool ms per level', pl, 'parity n/a')
" | tee -a gpurun_out/r2bt_ab.txt
  tail -1 gpurun_out/r2bt.err | cut -c1-200
done
