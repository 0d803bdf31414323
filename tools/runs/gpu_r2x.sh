set -x
cd $GRAFT_REPO_ROOT
nproc; lscpu | grep -i "model name\|^CPU(s)"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 900 python bench.py --steps 12 --warmup 3 --no-extra > gpurun_out/r2x_bench.json 2> gpurun_out/r2x.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2x_bench.json"))
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"])
print(json.dumps(d["e2e"], indent=1))
PY
tail -3 gpurun_out/r2x.err
