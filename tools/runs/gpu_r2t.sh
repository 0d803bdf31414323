set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2t_bench.json"))
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"])
r=d["roofline"]; print(r["frac"], r["deepest_launch"], r["whole_step"])
print(d.get("e2e_entry")); print(d.get("parity",{}).get("status"))
x=d["extra"]; print(json.dumps(x)[:3000])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
tail -3 gpurun_out/r2t.err
