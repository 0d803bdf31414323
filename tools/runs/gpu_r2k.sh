set -x
cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --steps 20 --warmup 3 --streams $1 --heads $2 --no-extra --no-cpu-baseline $3 2>> gpurun_out/r2k.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('streams $1 heads $2 $3 plain=$AZN_HEADS_PLAIN', round(d['value']), round(d['ms_per_step'],4))"; }
run 1 mma
AZN_HEADS_PLAIN=1 run 1 mma
run 1 mma --no-pdl
run 1 gemm --no-pdl
AZN_HEADS_PLAIN=1 run 4 mma
run 4 mma
run 4 gemm
tail -3 gpurun_out/r2k.err
