cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -6
timeout 300 python tools/backbone_bench.py 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['frac_of_sustained_bf16'])
for l in d['layers'][:6]: print(l)
"
AZN_CONV1_PATCHES=1 timeout 300 python tools/backbone_bench.py 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('patches route', d['value'], d['ms_per_step'])
"
