cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/r2bi_bench.json 2> gpurun_out/r2bi.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2bi_bench.json"))
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"])
r=d["roofline"]; print(r["frac"], r["deepest_launch"]["frac"], r["whole_step"]["frac"])
print({k:v for k,v in d["e2e"].items() if k in ("value","ms_per_step","route_timing")}, d["e2e"]["bf16_nhwc_input"]["value"])
e=d.get("e2e_entry"); print(e["value"], e["seconds"], e["host_seconds"], e["backbone_only_images_per_s"]); print(d.get("parity",{}).get("status"))
x=d["extra"]
for r in x["nms"]["rows"]: print("nms", r["N"], r["thresh"], r["ms"], round(r["boxes_per_s"]/1e6,1), r["match"])
for r in x["roi_pool"]["rows"]: print("pool", r["variant"], r["map"], r["R"], r["ms"], r["frac"])
print(x["config2_default_cfg_map"]); print(x["config3_detection"]); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
PY
timeout 300 python tools/backbone_bench.py 2>/dev/null | tail -1 > gpurun_out/r2bi_backbone.json; python -c "
import json; d=json.load(open('gpurun_out/r2bi_backbone.json')); print(d['value'], d['ms_per_step'], d['frac_of_sustained_bf16'], d.get('cudnn_library_reference'))"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2bi_bench_reference.json; cut -c1-200 gpurun_out/r2bi_bench_reference.json
tail -2 gpurun_out/r2bi.err
