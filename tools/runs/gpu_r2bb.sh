#!/bin/bash
# e2e call: split upload route (some images raw f32, the others narrowed on the host meanwhile).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py -x -q -k "pipeline" 2>&1 | tail -4
for hn in auto off on split; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --host-narrow $hn 2>gpurun_out/r2bb_$hn.err | tail -1 > gpurun_out/r2bb_bench_$hn.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2bb_bench_$hn.json"))
e=d["e2e"]
print("$hn", round(d["value"]), "e2e", round(e["value"]), round(e["ms_per_step"],3), e["h2d_bytes_per_step"], e.get("raw_f32_images_per_batch"), e["route_timing"], (e.get("f32_upload") or {}).get("value"))
PY
done
