set -x
cd $GRAFT_REPO_ROOT
cp aznet_b200/libaznet_b200.so /tmp/lib_backup.so
AZN_NVCC_EXTRA="-DAZN_POOL_TRACE" python -c "
from aznet_b200 import _lib; _lib.build(force=True)" 2>&1 | tail -2
python - <<'PY' 2>&1 | grep -v "^$" | tail -40
import torch, sys
sys.path.insert(0, ".")
from aznet_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
feat = torch.from_numpy(synth.make_conv_maps(1, 512, 38, 63, seed=7)).to(dev).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
for R in (2000, 20000):
    rois = torch.from_numpy(synth.make_rois(R, 600, 1000, seed=3)).to(dev)
    out = torch.empty((R, 7, 7, 512), dtype=torch.bfloat16, device=dev)
    print("R", R, flush=True)
    for it in range(2):
        ops.roi_pool(feat, rois, layout="NHWC", out=out)
        torch.cuda.synchronize()
PY
cp /tmp/lib_backup.so aznet_b200/libaznet_b200.so
