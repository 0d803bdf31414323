set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --streams 1 --no-extra --no-cpu-baseline > gpurun_out/r2f_bench_s1.json 2> gpurun_out/r2f.err; cut -c1-300 gpurun_out/r2f_bench_s1.json
timeout 300 python bench.py --steps 10 --warmup 3 --streams 2 --no-extra --no-cpu-baseline > gpurun_out/r2f_bench_s2.json 2>> gpurun_out/r2f.err; cut -c1-300 gpurun_out/r2f_bench_s2.json
timeout 300 python bench.py --steps 40 --warmup 3 --streams 2 --no-extra --no-cpu-baseline > gpurun_out/r2f_bench_s2_k40.json 2>> gpurun_out/r2f.err; cut -c1-300 gpurun_out/r2f_bench_s2_k40.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2>> gpurun_out/r2f.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d.get("e2e_entry")); print(d.get("parity",{}).get("status"))
PY
tail -5 gpurun_out/r2f.err
