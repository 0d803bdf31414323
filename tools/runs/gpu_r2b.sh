set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -40
for m in 12 22; do
  python tools/microbench.py --only roi_pool --pool-mode $m --sizes 2000,20000 >> gpurun_out/r2b_pool_mode$m.jsonl 2>&1
  python tools/microbench.py --only roi_pool --pool-mode $m --sizes 2000,20000 --hw 30,50 >> gpurun_out/r2b_pool_mode$m.jsonl 2>&1
done
grep -h nhwc gpurun_out/r2b_pool_mode*.jsonl | cut -c1-260
