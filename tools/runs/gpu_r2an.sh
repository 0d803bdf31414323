cd $GRAFT_REPO_ROOT
N=2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 400 $T bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r2an.err | tail -1 > gpurun_out/r2an_n2.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2an_n2.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["clocks"])
print({k:v for k,v in d["e2e"].items() if k in ("value","ms_per_step","route_timing","route")})
print(d.get("e2e_entry",{}).get("value"), d.get("parity",{}).get("status"))
PY
tail -3 gpurun_out/r2an.err
timeout -k 10 200 $T bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>>gpurun_out/r2an.err | tail -1 | cut -c1-300
