set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_golden.py -m gpu -q -k nms_20000 2>&1 | tail -3
python - <<'PY'
import torch
dev=torch.device("cuda:0")
x=torch.empty(1<<30,dtype=torch.uint8,device=dev); y=torch.empty_like(x)
def t(fn,n=10):
    fn(); torch.cuda.synchronize(); best=1e9
    for _ in range(n):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
print("memset 1GiB GB/s", (1<<30)/t(lambda: x.zero_())/1e6)
print("copy 1GiB r+w GB/s", 2*(1<<30)/t(lambda: y.copy_(x))/1e6)
print("read-only (sum) GB/s", (1<<30)/t(lambda: x.view(torch.int32).sum())/1e6)
PY
ncu --set full --clock-control none --import-source on -k regex:roi_pool_keys_kernel -s 3 -c 1 -o gpurun_out/r2c_pool_v2 python tools/microbench.py --only roi_pool --pool-mode 22 --sizes 20000 > gpurun_out/r2c_ncu_v2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:roi_pool_keys_kernel -s 3 -c 1 -o gpurun_out/r2c_pool_v1 python tools/microbench.py --only roi_pool --pool-mode 12 --sizes 20000 > gpurun_out/r2c_ncu_v1.log 2>&1
ls -la gpurun_out/*.ncu-rep
