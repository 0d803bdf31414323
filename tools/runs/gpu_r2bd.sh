#!/bin/bash
# conv row re-use (RU kernels): parity of the backbone tests with / without the descriptor base offset, backbone bench A/B.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for r in 1 2 0; do
  echo "== AZN_CONV_REUSE=$r"
  AZN_CONV_REUSE=$r timeout 300 python -m pytest tests/test_gpu_backbone.py -q 2>&1 | tail -6
  AZN_CONV_REUSE=$r timeout 300 python tools/backbone_bench.py --no-cudnn 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],3), round(d['frac_of_sustained_bf16'],3), [(l['layer'], l['ms']) for l in d['layers'] if l['layer'].startswith('conv') and l['layer'] <= 'conv3_1'])"
done
