cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ah_bench.json"))
print(d["value"], d["ms_per_step"], d["clocks"])
print({k:v for k,v in d["e2e"].items() if k in ("value","ms_per_step","route_timing")})
print(d.get("e2e_entry"))
PY
tail -3 gpurun_out/r2ah.err
