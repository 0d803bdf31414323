"""Kernel timeline of test_proposals (batched route) under torch.profiler: total GPU busy time against the wall clock."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import benchlib as BL
from aznet_b200 import _lib, backbone, engine, net, ops, synth
from aznet_b200.detect import config as C, test as T
_lib.build(); _lib.require_device()
dev = torch.device("cuda:0")
head = engine.AZHeadWeights(synth.make_az_weights(seed=3, zoom_bias=0.1), dev)
bw = backbone.make_vgg16_weights(seed=5)
bw["conv1_1"] = (bw["conv1_1"][0] / np.float32(128.0), bw["conv1_1"][1])
bb = backbone.VGG16Backbone(bw, dev)
nets = {"full": net.Net(head, "az", backbone=bb, name="az_vgg16"), "fc": net.Net(head, "az", name="az_vgg16")}
base = synth.make_images(32, 600, 1000, seed=1000)
images = [base[i % 32] for i in range(256)]
with tempfile.TemporaryDirectory() as tmp:
    C.cfg.TEST.MAX_SIZE, C.cfg.SEAR.BATCH_SIZE, C.cfg.ROOT_DIR = 800, 1000, tmp
    C.cfg_set_path("probe")
    C.cfg_set_mode("Test", 0.07)
    C.cfg.SEAR.NUM_PROPOSALS = 300
    with BL._Quiet():
        T.test_proposals(nets, synth.InMemoryImdb(images[:64], num_classes=21, name="warm"))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        t0 = time.perf_counter()
        with BL._Quiet():
            T.test_proposals(nets, synth.InMemoryImdb(images, num_classes=21, name="probe"))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    busy = sum(e.device_time for e in ev) / 1e3
    print("wall ms", round(dt * 1e3, 1), "sum of GPU activity ms", round(busy, 1), "batches", 4)
    agg = {}
    for e in ev:
        k = e.name[:70]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.device_time / 1e3
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        print("%8.2f ms %5d  %s" % (ms, n, k))
