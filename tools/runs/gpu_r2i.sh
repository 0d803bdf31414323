set -x
cd $GRAFT_REPO_ROOT
for cfg in "1 gemm" "1 mma" "2 gemm" "2 mma" "3 mma" "4 mma" "1 mma" "1 gemm"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 3 --streams $1 --heads $2 --no-extra --no-cpu-baseline 2>> gpurun_out/r2i.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('streams $1 heads $2', round(d['value']), round(d['ms_per_step'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
python - <<'PY'
import sys, json, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import bench, benchlib as BL
from aznet_b200 import engine, synth
dev = torch.device("cuda:0")
head = engine.AZHeadWeights(synth.make_az_weights(seed=3, zoom_bias=bench.ZOOM_BIAS), dev)
for k in range(2):
    print(json.dumps(BL.entry_point_throughput(dev, head, bench.CFG)))
PY
tail -3 gpurun_out/r2i.err
