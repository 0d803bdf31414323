set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
