cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys, json, torch
sys.path.insert(0, "tools")
from aznet_b200 import _lib, ops, synth
import benchlib as BL
dev = torch.device("cuda:0")
for mode in (0, 64, 0, 64):
    _lib.lib().azn_nms_tune(mode)
    for N in (8000, 20000):
        d = torch.from_numpy(synth.make_dets(N, seed=3)).to(dev)
        for th in (0.3, 0.7):
            best, mean = BL.time_best(lambda: ops.nms(d, th), dev, iters=10)
            print(mode, N, th, round(best, 4), round(N / best / 1e3, 1), "M boxes/s")
_lib.lib().azn_nms_tune(0)
PY
