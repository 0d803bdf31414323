"""Throughput through detect.test.test_proposals (the e2e_entry block of bench.py) as a stand-alone tool:
    python tools/probes/entry_chunk.py                                   one GPU (AZN_BACKBONE_CHUNK=16|32: A/B of the chunk size)
    torchrun --nproc-per-node N tools/probes/entry_chunk.py [images]     N GPUs: test_proposals shards the image database over the
                                                                         ranks itself; rank 0 prints the whole job's images/s"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
import bench
import benchlib as BL
from aznet_b200 import _lib, engine, synth
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
_lib.build(); _lib.require_device()
head = engine.AZHeadWeights(synth.make_az_weights(seed=3, zoom_bias=bench.ZOOM_BIAS), dev)
n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 1024 * world
out = BL.entry_point_throughput(dev, head, bench.CFG, n_images=n_images)
if world > 1:
    t = torch.tensor([out["seconds"]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["seconds"] = round(float(t.item()), 4)
    out["value"] = round(n_images / out["seconds"], 1)
    out["n_gpus"] = world
if rank == 0:
    print(os.environ.get("AZN_BACKBONE_CHUNK"), json.dumps(out))
if world > 1:
    dist.destroy_process_group()
