set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2m_bench.json"))
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"])
print({k:v for k,v in d["e2e"].items() if k!="bf16_nhwc_input"}); print(d["e2e"]["bf16_nhwc_input"])
r=d["roofline"]; print(r["frac"], r["deepest_launch"], r["whole_step"], r["share_of_step"])
print(d.get("e2e_entry")); print(d.get("parity",{}).get("status")); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
x=d["extra"]; print(x["config2_default_cfg_map"], x["config3_detection"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_bench_reference.json 2>> gpurun_out/r2m.err; cut -c1-200 gpurun_out/r2m_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
tail -3 gpurun_out/r2m.err
