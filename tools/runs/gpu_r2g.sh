set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 3 --streams 2 --heads gemm --no-extra --no-cpu-baseline > gpurun_out/r2g_bench_gemmheads.json 2> gpurun_out/r2g.err; cut -c1-260 gpurun_out/r2g_bench_gemmheads.json
timeout 300 python bench.py --steps 20 --warmup 3 --streams 2 --heads mma --no-extra --no-cpu-baseline > gpurun_out/r2g_bench_mmaheads.json 2>> gpurun_out/r2g.err; cut -c1-260 gpurun_out/r2g_bench_mmaheads.json
timeout 300 python bench.py --steps 20 --warmup 3 --streams 1 --heads mma --no-extra --no-cpu-baseline > gpurun_out/r2g_bench_mmaheads_s1.json 2>> gpurun_out/r2g.err; cut -c1-260 gpurun_out/r2g_bench_mmaheads_s1.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2>> gpurun_out/r2g.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2g_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]); print(d.get("e2e_entry")); print(d.get("parity",{}).get("status"))
for r in d["roofline"]["per_level"]:
    if r["stage"] in ("heads","int7"): print(r)
PY
tail -5 gpurun_out/r2g.err
