cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_backbone.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_skip.py tests/test_gpu_zz_concurrency.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/backbone_bench.py --no-cudnn 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['frac_of_sustained_bf16'])
"
AZN_NO_POOL_FUSION=1 timeout 300 python tools/backbone_bench.py --no-cudnn 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('separate pools', d['value'], d['ms_per_step'])
"
