#!/bin/bash
# Re-entry check of round 2 at HEAD (container re-created): full GPU suite, smoke, the default bench (wall clock noted),
# reference arm, backbone.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
t1=$(date +%s); echo "tests: $((t1-t0)) s"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
t2=$(date +%s); echo "smoke: $((t2-t1)) s"
timeout 600 python bench.py > gpurun_out/r2bm_bench.json 2> gpurun_out/r2bm.err
t3=$(date +%s); echo "bench (defaults): $((t3-t2)) s"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2bm_bench.json"))
print(d["value"], d["ms_per_step"], d["steps"], d["warmup"], d["gpu_launches"], d["clocks"])
r=d["roofline"]; print(r["frac"], r["deepest_launch"]["frac"], r["whole_step"]["frac"])
print({k:v for k,v in d["e2e"].items() if k in ("value","ms_per_step")}, d["e2e"].get("route_timing",{}).get("chosen"), d["e2e"]["bf16_nhwc_input"]["value"])
e=d.get("e2e_entry"); print(e["value"], e["seconds"], e["backbone_only_images_per_s"]); print(d.get("parity",{}).get("status"))
x=d["extra"]
for r in x["nms"]["rows"]: print("nms", r["N"], r["thresh"], r["ms"], round(r["boxes_per_s"]/1e6,1), r["match"])
for r in x["roi_pool"]["rows"]: print("pool", r["variant"], r["map"], r["R"], r["ms"], r["frac"])
print(x["config2_default_cfg_map"]); print(x["config3_detection"]); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2bm_bench_reference.json; cut -c1-300 gpurun_out/r2bm_bench_reference.json
t4=$(date +%s); echo "reference arm: $((t4-t3)) s"
timeout 200 python tools/backbone_bench.py --no-cudnn 2>/dev/null | tail -1 > gpurun_out/r2bm_backbone.json; python -c "
import json; d=json.load(open('gpurun_out/r2bm_backbone.json')); print(d['value'], d['ms_per_step'], d['frac_of_sustained_bf16'])"
tail -2 gpurun_out/r2bm.err
