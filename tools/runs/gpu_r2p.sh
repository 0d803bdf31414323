set -x
cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:"fc_gemm_kernel|az_heads_kernel" -s 36 -c 12 -o gpurun_out/r2p_gemm python bench.py --steps 2 --warmup 3 --no-graph --no-extra --no-cpu-baseline > gpurun_out/r2p_ncu.log 2>&1
ls -la gpurun_out/r2p_gemm.ncu-rep
