cd $GRAFT_REPO_ROOT
for b in 16 32 64; do timeout 300 python tools/backbone_bench.py --no-cudnn --batch $b 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['batch'], round(d['value'],1), round(d['ms_per_step'],3), round(d['frac_of_sustained_bf16'],3), [ (l['layer'], l['ms']) for l in d['layers'] if l['layer'] in ('conv1_2','conv4_2','conv5_1','conv5_3')])
"; done
