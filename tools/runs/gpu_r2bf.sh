#!/bin/bash
cd $GRAFT_REPO_ROOT
for c in 256 512 256 512; do
  AZN_CONV_BN128_CIN=$c timeout 300 python tools/backbone_bench.py --no-cudnn 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('cin<=$c', round(d['value'],1), round(d['ms_per_step'],3), round(sum(l['ms'] for l in d['layers']),3))"
done
AZN_CONV_BN128_CIN=512 timeout 300 python -m pytest tests/test_gpu_backbone.py -q 2>&1 | tail -3
for b in 16 32; do AZN_CONV_BN128_CIN=512 timeout 300 python tools/backbone_bench.py --no-cudnn --batch $b 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('batch $b', round(d['value'],1), round(d['ms_per_step'],3))"; done
