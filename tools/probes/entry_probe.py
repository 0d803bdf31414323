"""Where does e2e_entry lose time?  test_proposals throughput in a fresh process, after an azn_nms call (green-context
partitions exist), and after the host-narrowing worker pool has been used."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import benchlib as BL
from aznet_b200 import _lib, engine, ops, synth
_lib.build(); _lib.require_device()
dev = torch.device("cuda:0")
CFG = dict(scales=(600,), max_size=800, min_side=10, tz=0.5, num_proposals=300, batch_size=1000, dedup=1. / 16., eps=1e-14)
head = engine.AZHeadWeights(synth.make_az_weights(seed=3, zoom_bias=0.1), dev)
def run(tag):
    r = BL.entry_point_throughput(dev, head, CFG)
    print(tag, r["value"], r["seconds"], r["host_seconds"], r["backbone_only_images_per_s"], flush=True)
run("fresh")
run("again")
x = torch.randn(64 * 512 * 30 * 50)
o = torch.empty(x.numel(), dtype=torch.bfloat16)
ops.host_f32_to_bf16(x, o, 16)
run("after host pool")
d = torch.from_numpy(synth.make_dets(20000, seed=3)).to(dev)
ops.nms(d, 0.5); torch.cuda.synchronize()
run("after nms (partitions)")
