set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -q -x -k "roi_pool" 2>&1 | tail -15
for m in 22 322; do timeout 300 python tools/microbench.py --only roi_pool --pool-mode $m --sizes 2000,8000,20000 --hw 38,63 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['pool_mode'], d['variant'], d['map'], d['R'], round(d['ms_best'],4), round(d['gbs']), round(d['frac_of_measured_hbm'],3))
"; done
