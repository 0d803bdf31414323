#!/bin/bash
# conv RU with 256 x 128 tiles for the wide layers (A/B by input width)
cd $GRAFT_REPO_ROOT
for c in 0 128 256 512; do
  AZN_CONV_BN128_CIN=$c timeout 300 python tools/backbone_bench.py --no-cudnn 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('cin<=$c', round(d['value'],1), round(d['ms_per_step'],3), [(l['layer'], l['ms']) for l in d['layers'] if l['layer'] >= 'conv3_1' and l['layer'].startswith('conv')])"
done
