"""A/B of the backbone chunk size behind test_proposals (AZN_BACKBONE_CHUNK): prints the e2e_entry block."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
import benchlib as BL
from aznet_b200 import _lib, engine, synth
_lib.build(); _lib.require_device()
dev = torch.device("cuda:0")
head = engine.AZHeadWeights(synth.make_az_weights(seed=3, zoom_bias=bench.ZOOM_BIAS), dev)
print(os.environ.get("AZN_BACKBONE_CHUNK"), json.dumps(BL.entry_point_throughput(dev, head, bench.CFG)))
