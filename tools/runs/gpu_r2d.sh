set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 3000 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-coop --no-extra --no-cpu-baseline > gpurun_out/r2d_bench_nocoop.json 2>> gpurun_out/r2d_bench.err; cut -c1-400 gpurun_out/r2d_bench_nocoop.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2d_bench_reference.json 2>> gpurun_out/r2d_bench.err; cut -c1-600 gpurun_out/r2d_bench_reference.json
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
