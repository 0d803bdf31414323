#!/bin/bash
# Detection step (config #3): ROI-pool kernel A/B -- staged (default) vs the direct kernels.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for pool in staged per_roi direct staged; do
  timeout 200 python tools/detbench.py --pool $pool --cpu-images 0 2>gpurun_out/r2bv.err | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
print('pool $pool', round(d['value']), round(d['ms_per_step'], 4), {k: v for k, v in d.items() if k in ('stages_ms', 'per_stage_ms')} or [k for k in d.keys()])
" | tee -a gpurun_out/r2bv_ab.txt
  tail -1 gpurun_out/r2bv.err | cut -c1-200
done
