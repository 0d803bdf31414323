set -x
cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --steps 24 --warmup 3 --streams $1 --no-extra --no-cpu-baseline 2>> gpurun_out/r2l.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('streams $1 mid=$AZN_FC_MID_TILE', round(d['value']), round(d['ms_per_step'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
run 4
run 6
run 8
run 1
AZN_FC_MID_TILE=2128 run 4
AZN_FC_MID_TILE=1256 run 4
AZN_FC_MID_TILE=2256 run 4
AZN_FC_MID_TILE=2128 run 1
AZN_FC_MID_TILE=1256 run 1
AZN_FC_MID_TILE=2256 run 1
grep -v "^import\|^d=json" gpurun_out/r2l.err | tail -3
