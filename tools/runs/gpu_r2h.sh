set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_search.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --streams 2 --heads mma --no-extra --no-cpu-baseline > gpurun_out/r2h_bench_s2.json 2> gpurun_out/r2h.err; cut -c1-260 gpurun_out/r2h_bench_s2.json
timeout 300 python bench.py --steps 20 --warmup 3 --streams 1 --heads mma --no-extra --no-cpu-baseline > gpurun_out/r2h_bench_s1.json 2>> gpurun_out/r2h.err; cut -c1-260 gpurun_out/r2h_bench_s1.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench_s1.json"))
for r in d["roofline"]["per_level"]:
    if r["stage"] in ("heads",): print(r)
PY
tail -3 gpurun_out/r2h.err
