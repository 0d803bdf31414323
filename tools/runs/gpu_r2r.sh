set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_detect.py -m gpu -q -x -k "roi_pool or detect" 2>&1 | tail -4
for hw in 38,63 30,50; do
python tools/microbench.py --only roi_pool --pool-mode 22 --sizes 2000,8000,20000 --hw $hw 2>&1 | grep nhwc | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['variant'], d['map'], d['R'], round(d['ms_best'],4), round(d['gbs']), round(d['frac_of_measured_hbm'],3))"
done
