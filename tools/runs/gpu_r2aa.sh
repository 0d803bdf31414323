set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "nms" 2>&1 | tail -8
timeout 300 python tools/microbench.py --only nms --sizes 2000,8000,20000 --nms-phases 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['N'], d['thresh'], round(d['ms_best'],4), round(d['boxes_per_s']/1e6,1), d.get('phases'))
"
