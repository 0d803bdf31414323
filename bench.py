#!/usr/bin/env python
"""bench.py -- AZ-Net proposal throughput (images/sec) on B200, the metric of BASELINE.json.

A step = one pass of the adaptive search (all levels: ROI max-pool over the cached conv5_3 map,
fc heads, decode/clip, zoom-subdivide, dedup, top-N) over one batch of 64 synthetic 600x1000 images
per GPU, AZ-Net VGG16 head widths, PASCAL config (experiments/cfgs/voc.yml: MAX_SIZE 800 -> conv5_3
512x30x50, BATCH_SIZE 1000, 300 proposals).  The conv5_3 maps are the input of the path (the backbone
runs once per image before the search and is out of scope).

  value   whole-job images/sec with the maps resident in HBM (NHWC bf16)
  e2e     the same through the host-facing call: f32 NCHW maps in pinned host memory -> H2D -> layout
          conversion -> search -> D2H of the proposal lists, every step
  roofline  the int6 GEMM (25088 -> 4096), time-weighted over its launches of a step, vs the measured bf16 peak
  parity    oracle vs CUDA path on the first images of the timed batch (outside the timed region)
  extra     the other metrics of BASELINE.json in the same record: ROI-pool GB/s, NMS boxes/s (config #4), config #3,
            config #2 on the default-cfg 38x63 map
  e2e_entry images/s THROUGH detect.test.test_proposals (the call tools/prop_az.py makes), backbone included
  cpu_baseline / --impl reference  the reference's CPU path on the host cores: its own lib/detect/test.py control
            flow + compiled div.pyx + its own layer sources (oracle/_ref) when present, else the oracle port

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--job IMAGES] [--no-extra]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IM_H, IM_W, BATCH = 600, 1000, 64
CFG = dict(scales=(600,), max_size=800, min_side=10, tz=0.5, num_proposals=300, batch_size=1000,
           dedup=1. / 16., eps=1e-14)
ZOOM_BIAS = 0.1          # ~40 % of regions zoom at Tz = 0.5 with the seed-3 He-init heads
WORKLOAD = ("AZ-Net VGG16 PASCAL config (voc.yml: MAX_SIZE 800 -> conv5_3 512x30x50), batch of 64 synthetic "
            "600x1000 images per GPU, search from cached conv5_3, Tz=0.5, 300 proposals")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md).  NVML polled every millisecond
    from a thread (a 10-step timed region lasts ~10 ms: `nvidia-smi -lms` cannot land a sample in it reliably);
    nvidia-smi is the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nv, self.h, self.samples, self.stop_flag, self.thread = None, None, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            nv = self.nv
            try:
                mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            except Exception:
                mx = None
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            reasons = sorted({k for _, rs in self.samples for k, b in bits.items() if rs & b})
            sm = [float(c) for c, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm),
                    "source": "nvml, 1 ms polling over the timed regions"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 9 for k in range(4) if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}


class OracleRunner:
    """The reference's CPU path on images of the bench workload.

    kind "reference": the reference's OWN code wherever it can run here -- lib/detect/test.py::im_propose (converted
    py2->py3, oracle/_ref/pyref), lib/utils/div.pyx compiled unmodified, and Forward_cpu of its own roi_pooling /
    relu / sigmoid layer sources (oracle/_ref/libcaffe_layers_ref.so); the InnerProduct layers' sgemm goes to
    torch-CPU (MKL) on `threads` cores because the reference's BLAS is un-vendored.  Like the reference it redoes the
    host image blob (mean-subtract + cv2.resize) at every level (test.py:208) and runs ROI pooling and the control
    flow on one core.  The backbone is out of scope on both arms: the 'full' net hands back the cached conv5_3 map.
    kind "port": the oracle restatement (oracle/az_oracle.py), when oracle/_ref is not populated."""

    def __init__(self, weights, threads, n_maps=8, seed=7):
        from aznet_b200 import synth
        from oracle import az_oracle as O
        from oracle import build_ref, ref_caffe
        self.O = O
        self.cfg = O.OracleCfg(TEST_MAX_SIZE=CFG["max_size"], Tz=CFG["tz"], NUM_PROPOSALS=CFG["num_proposals"],
                               BATCH_SIZE=CFG["batch_size"])
        s = O.im_scale_for((IM_H, IM_W), self.cfg)[0]
        fh, fw = synth.conv_shape(IM_H, IM_W, s)
        self.conv = synth.make_conv_maps(n_maps, 512, fh, fw, seed=seed)
        self.regions = 0
        self.images = 0
        self.threads = threads
        ref = build_ref.load_pyref() if ref_caffe.available() else None
        if ref is not None:
            self.kind = "reference"
            self.rtest, rconfig = ref[0], ref[1]
            rc = rconfig.cfg
            rc.TEST.MAX_SIZE, rc.SEAR.BATCH_SIZE = CFG["max_size"], CFG["batch_size"]
            rconfig.cfg_set_mode("Test", CFG["tz"])
            rc.SEAR.NUM_PROPOSALS = CFG["num_proposals"]
            self.cur = None
            self.full = O.OracleNet(weights, "az", cfg=self.cfg, threads=threads, layers="ref", backbone=lambda data: self.cur)
            self.fc = O.OracleNet(weights, "az", cfg=self.cfg, threads=threads, layers="ref")
            self.im = np.zeros((IM_H, IM_W, 3), np.uint8)
            self.what = ("the reference's own lib/detect/test.py::im_propose + compiled div.pyx + its own ROIPooling/ReLU/Sigmoid "
                         "layer sources (oracle/_ref); sgemm heads through torch-CPU on %d threads; ROI-pool, host image blob "
                         "per level and control flow on 1 core like the reference" % threads)
        else:
            self.kind = "port"
            self.net = O.OracleNet(weights, "az", cfg=self.cfg, threads=threads)
            self.what = ("oracle port of lib/detect/test.py + Caffe layers (oracle/_ref not populated); ROI-pool and control "
                         "flow single-threaded like the reference, sgemm heads on %d threads" % threads)

    def run(self, n_images, start=0):
        import contextlib
        import io
        import re
        for i in range(n_images):
            j = (start + i) % self.conv.shape[0]
            if self.kind == "reference":
                self.cur = self.conv[j:j + 1]
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    self.rtest.im_propose({"full": self.full, "fc": self.fc}, self.im)
                self.regions += int(re.search(r"evaluate (\d+) regions", buf.getvalue()).group(1))
            else:
                nets = {"full": self.net, "fc": self.net}
                _, _, info = self.O.im_propose(nets, (IM_H, IM_W, 3), self.cfg, conv={"conv5_3": self.conv[j:j + 1]},
                                               return_scores=True)
                self.regions += info["num_eval"]
            self.images += 1


def reference_entry_point(weights, threads, n_images=2):
    """The reference-arm twin of `e2e_entry`: the reference's own test_proposals(net, imdb) (oracle/_ref/pyref) over PNG
    files, its 'full' net = VGG16 conv1_1..conv5_3 in fp32 on the host cores (torch-CPU conv2d: the class of library
    Caffe's im2col + sgemm is) + the AZ head through the reference's layer sources.  None when oracle/_ref is absent."""
    import contextlib
    import io
    import tempfile
    import cv2
    from aznet_b200 import backbone, synth
    from oracle import az_oracle as O
    from oracle import build_ref, ref_caffe
    ref = build_ref.load_pyref() if ref_caffe.available() else None
    if ref is None:
        return None
    rtest, rconfig = ref[0], ref[1]
    rc = rconfig.cfg
    ocfg = O.OracleCfg(TEST_MAX_SIZE=CFG["max_size"], Tz=CFG["tz"], NUM_PROPOSALS=CFG["num_proposals"], BATCH_SIZE=CFG["batch_size"])
    bw = backbone.make_vgg16_weights(seed=5)
    bw["conv1_1"] = (bw["conv1_1"][0] / np.float32(128.0), bw["conv1_1"][1])     # as tools/benchlib.py::entry_point_throughput
    full = O.OracleNet(weights, "az", cfg=ocfg, threads=threads, layers="ref", name="az_vgg16",
                       backbone=lambda data: O.vgg16_conv5(bw, data, threads=threads))
    fc = O.OracleNet(weights, "az", cfg=ocfg, threads=threads, layers="ref", name="az_vgg16")
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i, im in enumerate(synth.make_images(n_images, IM_H, IM_W, seed=1000)):
            paths.append(os.path.join(tmp, "%d.png" % i))
            cv2.imwrite(paths[-1], im)
        imdb = synth.SyntheticImdb(paths, num_classes=21, name="bench_ref")
        rc.TEST.MAX_SIZE, rc.SEAR.BATCH_SIZE, rc.ROOT_DIR = CFG["max_size"], CFG["batch_size"], tmp
        rconfig.cfg_set_path("bench")
        rconfig.cfg_set_mode("Test", CFG["tz"])
        rc.SEAR.NUM_PROPOSALS = CFG["num_proposals"]
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            rtest.test_proposals({"full": full, "fc": fc}, imdb)
        dt = time.perf_counter() - t0
    return {"value": n_images / dt, "unit": "images/s", "images": n_images, "seconds": dt, "cores": threads,
            "what": "the reference's own test_proposals over PNG files: cv2.imread, host image blob per level, VGG16 conv stack "
                    "in fp32 on the host cores, AZ head, proposals.pkl"}


def workload_config(world, job=0):
    """The `config` object of BOTH arms (equal for equal --gpus / --job): what is computed, not how a given arm ran it --
    an arm's own run details live beside it (`run` on the GPU arm, `cpu_baseline.sample` on the reference arm)."""
    cfg = {"workload": WORKLOAD, "global_batch": world * BATCH,
           "l2": "working set per step (216 MB bf16 weights + pooled rows) exceeds the 126 MB L2; inputs rotate over 2 distinct batches",
           "parallelism": "image-sharded x%d, no collective on the hot path" % world}
    if job:
        cfg["job"] = "BASELINE config #5: %d images = %d batches of 64 sharded over %d rank(s)" % (job, job // BATCH, world)
    return cfg


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, same config / metric /
    unit.  One step = a bounded sample of 4 images of the 64-image batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from aznet_b200 import synth
    threads = os.cpu_count() or 1
    w = synth.make_az_weights(seed=3, zoom_bias=ZOOM_BIAS)
    per_step = 4
    runner = OracleRunner(w, threads)
    for k in range(max(args.warmup, 1)):
        runner.run(1, k)
    runner.regions = runner.images = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        runner.run(per_step, k * per_step)
    dt = time.perf_counter() - t0
    n = per_step * args.steps
    ips = n / dt
    line = {
        "impl": "reference", "metric": "AZ proposal images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the GPU arm's `config`, key for key; what this arm actually timed -- a bounded sample of that workload -- is stated
        # in cpu_baseline.sample
        "config": workload_config(args.gpus, args.job),
        "regions_per_image": runner.regions / max(runner.images, 1),
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": runner.kind,
                         "sample": "a step = a bounded sample of %d images of the 64-image batch; %d images in %.1f s (%s)" % (
                             per_step, n, dt, runner.what)},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_extra:
        try:
            ent = reference_entry_point(w, threads)
            if ent is not None:
                line["e2e_entry"] = ent
        except Exception as e:                                        # the entry-point twin is a side figure
            line["e2e_entry"] = {"error": repr(e)}
    print(json.dumps(line))


def ncu_traffic():
    """DRAM bytes per launch of the roofline kernel.  It cannot be measured by the run itself (that needs ncu): the
    figure is STATIC, read from the committed summary of an `ncu --set full` capture (profiles/traffic.json names the
    capture and the commit it was taken at).  Returns (bytes or None, provenance string)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            k, v = next(iter(json.load(f).items()))
        return int(v["dram_bytes_per_launch"]), "static: %s (%s)" % (v.get("source", "profiles/traffic.json"), v.get("commit", "commit not recorded"))
    except Exception:
        return None, "no capture committed"


def parity_block(eng, dev_maps, host_maps, weights, n_check=4):
    """Oracle vs CUDA path on the first images of the timed batch, OUTSIDE the timed region: the fp32 oracle on the
    bf16-rounded weights and maps (the product's storage precision) with bf16 activation storage emulated."""
    import torch
    from oracle import az_oracle as O
    eng.propose(dev_maps)
    boxes, scores, n_eval, _ = eng.results()
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()
    wq = {k: (bf(v[0]), v[1]) for k, v in weights.items()}
    cfg = O.OracleCfg(TEST_MAX_SIZE=CFG["max_size"], Tz=CFG["tz"], NUM_PROPOSALS=CFG["num_proposals"], BATCH_SIZE=CFG["batch_size"])
    net = O.OracleNet(wq, "az", cfg=cfg, threads=os.cpu_count() or 1, act_round=O.round_bf16)

    def iou(a, b):
        x1, y1 = np.maximum(a[:, None, 0], b[None, :, 0]), np.maximum(a[:, None, 1], b[None, :, 1])
        x2, y2 = np.minimum(a[:, None, 2], b[None, :, 2]), np.minimum(a[:, None, 3], b[None, :, 3])
        inter = np.clip(x2 - x1 + 1, 0, None) * np.clip(y2 - y1 + 1, 0, None)
        aa = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
        ab = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        return inter / (aa[:, None] + ab[None, :] - inter)
    ne_g, ne_o, rec_o, rec_g, dtop = [], [], [], [], []
    hm = host_maps.numpy()
    for i in range(n_check):
        Y, sc, info = O.im_propose({"full": net, "fc": net}, (IM_H, IM_W, 3), cfg, conv={"conv5_3": bf(hm[i:i + 1])}, return_scores=True)
        m = iou(Y, boxes[i])
        ne_g.append(int(n_eval[i]))
        ne_o.append(int(info["num_eval"]))
        rec_o.append(float((m.max(1) >= 0.9).mean()))             # oracle proposals recovered by the CUDA path
        rec_g.append(float((m.max(0) >= 0.9).mean()))             # and vice versa
        top = min(20, len(sc), len(scores[i]))
        dtop.append(float(np.abs(np.sort(scores[i])[::-1][:top] - np.sort(sc)[::-1][:top]).max()))
    ok = all(abs(a - b) <= max(2, 0.03 * b) for a, b in zip(ne_g, ne_o)) and min(rec_o + rec_g) >= 0.95 and max(dtop) <= 3e-2
    return {"images": n_check, "n_eval_gpu": ne_g, "n_eval_oracle": ne_o, "recall_of_oracle_proposals_iou0.9": rec_o,
            "recall_of_gpu_proposals_iou0.9": rec_g, "top20_score_max_abs_diff": dtop,
            "tolerance": "regions evaluated within 3 % (+-2), recall >= 0.95 both ways at IoU 0.9, top-20 scores within 3e-2 "
                         "(bf16 operands and activations; tests/test_gpu_golden.py::test_full_width_search_vs_oracle)",
            "status": "green" if ok else "RED"}


def extra_blocks(dev, head, hbm_peak):
    """BASELINE.json's other metrics and configs, measured in the same process right after the headline (outside its
    timed region, bounded to a few seconds each): see tools/benchlib.py."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import benchlib as BL
    from oracle import az_oracle as O
    extra = {}
    wr = BL.hbm_write_only_gbs(dev)
    extra["roi_pool"] = {"metric": "ROI max-pool achieved HBM GB/s on the algorithmic bytes (map once + 20 B/ROI + pooled rows once)",
                         "peak_gbs": hbm_peak, "hbm_write_only_gbs": round(wr, 1),
                         "note": "frac = achieved / the measured COPY rate (read + write bytes); the kernel is >99 % writes, and a plain "
                                 "streaming-store kernel reaches hbm_write_only_gbs on this GPU",
                         "rows": BL.roi_pool_sweep(dev, hbm_peak, iters=5)}
    extra["nms"] = {"metric": "greedy NMS boxes/s (whole azn_nms call: sort + mask + chain + compaction)",
                    "rows": BL.nms_sweep(dev, oracle=O, iters=5)}
    extra["config2_default_cfg_map"] = BL.search_throughput(dev, head, dict(CFG, max_size=1000, batch_size=10000))
    extra["config3_detection"] = BL.detection_throughput(dev, head, CFG)
    entry = BL.entry_point_throughput(dev, head, CFG)
    return extra, entry


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying CUDA graphs")
    ap.add_argument("--no-pdl", action="store_true", help="plain stream order instead of programmatic dependent launch (A/B)")
    ap.add_argument("--no-coop", action="store_true", help="plain (PDL) launch of the persistent GEMM instead of the cooperative launch (A/B)")
    ap.add_argument("--streams", type=int, default=4, choices=(1, 2, 3, 4, 6, 8),
                    help="batches in flight for the device-resident figure: 2 = two engines on two CUDA streams, so that the latency-bound "
                         "kernels of one batch (ROI pool, decode / subdivide, selection) run next to the other batch's tensor-core GEMMs")
    ap.add_argument("--pool-tune", type=int, default=-1, help="A/B: azn_roi_pool_tune code for the whole run (500: direct kernel with one CTA per ROI)")
    ap.add_argument("--pool-per-roi-from", type=int, default=-1, help="A/B: first search level whose ROI pool runs one CTA per ROI (0: never; default: engine.POOL_PER_ROI_FROM_LEVEL)")
    ap.add_argument("--pool-staged-levels", default="", help="A/B: search levels (comma list, 1-based) whose ROI pool takes the staged kernel")
    ap.add_argument("--heads", default="mma", choices=("mma", "gemm"), help="output layers of the head: small mma.sync kernel or the persistent GEMM (A/B)")
    ap.add_argument("--host-narrow", default="auto", choices=("auto", "on", "off", "split"),
                    help="e2e call: round the f32 host maps to bf16 on the host cores before the upload; split: a third of the images cross "
                         "the link as f32 while the cores narrow the others (auto: time the routes on the first batch, keep the fastest)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra / e2e_entry / parity blocks (they run outside the timed regions)")
    ap.add_argument("--gather", default="peer", choices=("peer", "nccl"),
                    help="N > 1: how the ranks' proposal lists reach rank 0 -- peer: every step's lists are stored into a window of "
                         "rank 0's memory over NVLink by the step's last kernel, the job ends with a 4-byte all_reduce; nccl: one "
                         "all_gather of all lists at the end (round 1; also the fallback when the window cannot be mapped)")
    ap.add_argument("--job", type=int, default=0, help="strong-scaling mode (BASELINE config #5): a job of this many images "
                    "(a multiple of 64) sharded over the ranks in batches of 64; --steps is ignored")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from aznet_b200 import _lib, engine, ops, synth
    from aznet_b200.dist import CollectorGroup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.build()
    _lib.require_device()
    _lib.lib().azn_set_pdl(0 if args.no_pdl else 1)
    _lib.lib().azn_set_coop(0 if args.no_coop else 1)
    engine.HEADS_KERNEL = args.heads
    engine.POOL_STAGED_LEVELS = tuple(int(x) for x in args.pool_staged_levels.split(",") if x)
    if args.pool_per_roi_from >= 0:
        engine.POOL_PER_ROI_FROM_LEVEL = args.pool_per_roi_from
    if args.pool_tune >= 0:
        _lib.lib().azn_roi_pool_tune(args.pool_tune)
    _lib.lib().azn_az_heads_tune(1 if (args.no_graph or args.streams == 1) else 0)

    if args.job:
        assert args.job % (BATCH * world) == 0, "--job must be a multiple of 64 x the number of GPUs"
        args.steps = args.job // (BATCH * world)
    weights = synth.make_az_weights(seed=3, zoom_bias=ZOOM_BIAS)
    head = engine.AZHeadWeights(weights, dev)
    eng = engine.SearchEngine(head, BATCH, IM_H, IM_W, **CFG)
    fh, fw = synth.conv_shape(IM_H, IM_W, eng.scale)
    # distinct synthetic batches per rank (rotated, so consecutive steps never see the same maps): one per stream, >= 2
    n_sets = max(2, 1 if args.no_graph else args.streams)
    host_sets, dev_sets = [], []
    for sidx in range(n_sets):
        maps = synth.make_conv_maps(BATCH, 512, fh, fw, seed=7 + 1000 * rank + 100 * sidx)
        hp = torch.from_numpy(maps).pin_memory()
        host_sets.append(hp)
        dev_sets.append(ops.nchw_to_nhwc_bf16(hp.to(dev)))
    from aznet_b200.pipeline import ProposalPipeline
    # N > 1: every rank collects its batches' proposal lists on the device and the job does ONE NCCL all_gather
    # at the end of the K steps (inside the timed region) -- no collective per step, like test_proposals, which
    # appends per image and writes proposals.pkl once (lib/detect/test.py:508-539)
    # Two batches in flight (--streams 2): a second engine with its own buffers on a second stream.  Images are
    # independent, so consecutive batches have no dependency; the persistent GEMMs of the two streams take turns (they
    # are launched cooperatively: each needs every SM), while the small kernels of one batch fill the other batch's GEMM
    # time -- they use no tensor pipe and little shared memory, so they are co-resident with the GEMM's one CTA per SM.
    n_streams = 1 if args.no_graph else args.streams
    engines = [eng] + [engine.SearchEngine(head, BATCH, IM_H, IM_W, **CFG) for _ in range(n_streams - 1)]
    side = [torch.cuda.Stream(device=dev) for _ in range(n_streams)] if n_streams > 1 else []
    slots = (max(args.steps, args.warmup, 2) + n_streams - 1) // n_streams
    group, gather_route = None, None
    if world > 1:
        outs = [(e.out_boxes, e.out_scores, e.out_count) for e in engines]
        if args.gather == "peer":
            try:
                group, gather_route = CollectorGroup(slots, outs, peer=True), "peer"
            except RuntimeError as e:                 # collective decision: every rank raises or none does
                if rank == 0:
                    print("bench: %s -- falling back to the NCCL all_gather" % e, file=sys.stderr)
        if group is None:
            group, gather_route = CollectorGroup(slots, outs), "nccl"
    collectors = group.collectors if group else []
    collector = collectors[0] if collectors else None
    # the copy into the collection is the last kernel of the search (azn_collect_proposals, slot from a device-side
    # counter), so it is part of the replayed CUDA graph
    for e, c in zip(engines, collectors):
        e.collector = c
    after = None
    # --host-narrow auto: the pipeline times "upload the f32 batch" against "round it to bf16 on the host cores, chunk by
    # chunk, and upload half the bytes" on the first warm-up batch and keeps the faster route (aznet_b200/pipeline.py)
    pipe = ProposalPipeline(eng, tuple(host_sets[0].shape), depth=2, after_search=after, use_graph=not args.no_graph,
                            host_narrow=args.host_narrow)
    pipe_f32 = ProposalPipeline(eng, tuple(host_sets[0].shape), depth=2, use_graph=not args.no_graph) if args.host_narrow != "off" else None
    d2h_bytes = pipe.d2h_bytes
    # the same call with the host maps already in the engine's storage format (bf16 NHWC): half the PCIe bytes
    host_sets_bf16 = [d.cpu().pin_memory() for d in dev_sets]
    pipe_bf16 = ProposalPipeline(eng, tuple(host_sets[0].shape), depth=2, use_graph=not args.no_graph, layout="nhwc_bf16")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    graphs = []            # one CUDA graph per resident input set: the level loop's launch sequence is static

    def step_resident(i):
        if graphs and not eng.profile:
            g, n = graphs[i % n_sets]                 # set k is captured on engine k % n_streams
            if side:
                with torch.cuda.stream(side[i % n_streams]):
                    g.replay()
            else:
                g.replay()
            eng.launches += n
        else:
            eng.propose(dev_sets[i % n_sets])

    def fork():
        cur = torch.cuda.current_stream(dev)
        for st in side:
            st.wait_stream(cur)

    def join():
        cur = torch.cuda.current_stream(dev)
        for st in side:
            cur.wait_stream(st)

    pending = []

    def step_e2e(i):
        # host maps -> H2D (copy stream) -> layout conversion -> search -> D2H of the proposal lists; the upload
        # of step i overlaps the search of step i-1, every step moves its own h2d_bytes + d2h_bytes
        pending.append((pipe, pipe.submit(host_sets[i % n_sets])))
        if len(pending) == 2:
            p, t = pending.pop(0)
            p.result(t)

    def drain():
        while pending:
            p, t = pending.pop(0)
            p.result(t)

    def step_e2e_bf16(i):
        pending.append((pipe_bf16, pipe_bf16.submit(host_sets_bf16[i % n_sets])))
        if len(pending) == 2:
            p, t = pending.pop(0)
            p.result(t)

    def step_e2e_f32(i):
        pending.append((pipe_f32, pipe_f32.submit(host_sets[i % n_sets])))
        if len(pending) == 2:
            p, t = pending.pop(0)
            p.result(t)

    gathered = [None, None, None]
    gather_ms = [0.0]

    def timed(step_fn, steps, profile=False):
        barrier()
        for c in collectors:
            c.reset()
        eng.launches = 0
        eng.profile = profile
        eng.prof_events = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fork()
        for i in range(steps):
            step_fn(i)
        join()
        drain()                                       # e2e: the last proposals are on the host
        eg = torch.cuda.Event(enable_timing=True)
        eg.record()
        if world > 1:
            if gather_route == "peer" or (step_fn is step_resident and not profile):
                # the job's only exchange.  peer: the lists already sit in rank 0's window (every step's last kernel stored
                # them there), this is the closing fence; nccl: ONE all_gather of all ranks' proposal lists
                gathered[:] = group.gather()
            else:
                gathered[:] = [collector.gather(views=True)]
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        gather_ms[0] = eg.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(args.warmup):
        step_resident(i)
    if not args.no_graph:
        for k, e in enumerate(engines[1:], 1):
            e.propose(dev_sets[k])                    # warm-up of the other engines' buffers
        graphs.extend(engines[k % n_streams].capture(d) for k, d in enumerate(dev_sets))
        fork()
        for i in range(args.warmup):
            step_resident(i)
        join()
    for i in range(max(args.warmup, 2)):
        step_e2e(i)
    drain()
    for i in range(max(args.warmup, 2)):
        step_e2e_bf16(i)
    drain()
    if pipe_f32 is not None:
        for i in range(max(args.warmup, 2)):
            step_e2e_f32(i)
        drain()
    if group is not None:
        group.gather()                                # warm-up of the job's one exchange (buffers, NCCL channels)
        collector.gather(views=True)
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    launches = eng.launches
    gather_resident_ms = gather_ms[0]
    # per-kernel CUDA-event timings need host-issued launches: a second pass over the same K steps, not the timed one
    timed(step_resident, args.steps, profile=True)
    prof = eng.prof_summary()
    eng.profile = False
    regions = float(eng.n_eval.float().mean().item())
    if any(int(e.status.item()) != 0 for e in engines):
        raise RuntimeError("search capacity overflow during the bench")
    ms_e2e = timed(step_e2e, args.steps)
    ms_e2e_bf16 = timed(step_e2e_bf16, args.steps)
    narrowed = bool(pipe.narrow)
    if world > 1:                                     # the ranks decide for themselves; the extra timed pass must be collective
        t = torch.tensor([1.0 if narrowed else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        narrowed = bool(t.item() > 0)
    ms_e2e_f32 = timed(step_e2e_f32, args.steps) if (pipe_f32 is not None and narrowed) else None
    h2d_bytes = pipe.h2d_bytes
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        hbm_peak, tf_peak, which = peaks()
        value = world * BATCH * args.steps / (ms_total / 1e3)
        e2e = world * BATCH * args.steps / (ms_e2e / 1e3)
        top = prof["int6_deepest"]
        # int6 over ALL its launches of a step: algorithmic FLOPs of every level / the time of every level
        int6 = [r for r in prof["levels"] if r["stage"] == "int6"]
        k6, n6 = head.w6.shape[1], head.w6.shape[0]
        int6_ms = sum(r["ms"] for r in int6)
        int6_tf = sum(2.0 * r["m"] * n6 * k6 for r in int6) / (int6_ms * 1e-3) / 1e12 if int6_ms > 0 else 0.0
        step_flops = sum(r["m"] for r in int6) * 216119808.0          # FLOPs per ROI of the whole AZ head
        traffic, traffic_src = ncu_traffic()
        line = {
            "metric": "AZ proposal images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.job else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, args.job),
            "run": {"regions_per_image": regions,
                    "launch": "eager" if args.no_graph else "CUDA graph replay of the whole level loop (static launch sequence, device-side counts), "
                              "%d batch(es) in flight on %d stream(s); roofline events from a second, host-launched, single-stream pass over the same steps" % (n_streams, n_streams),
                    "lists_to_rank0": {
                        None: "single rank",
                        "peer": "every step's proposal lists are appended to rank 0's collection through a peer window (stores over NVLink by the "
                                "step's last kernel, azn_collect_proposals); the job ends with one 4-byte NCCL all_reduce as the fence, inside the timed region",
                        "nccl": "one NCCL all_gather of all K steps' proposal lists at the end, inside the timed region"}[gather_route]},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps, "h2d_gbs_per_rank": h2d_bytes / (ms_e2e / args.steps * 1e-3) / 1e9,
                    "input": "f32 NCHW conv5_3 maps in pinned host memory (the blobs the reference's 'fc' net is handed)",
                    "route": (("%d of the %d images cross the link as f32 while the host cores round the others to bf16, chunk by chunk "
                               "(azn_host_f32_to_bf16, %d threads), each narrowed chunk following the raw piece that was uploading meanwhile; "
                               "f32 / bf16 NCHW -> bf16 NHWC on the device" % (pipe.raw_images, BATCH, pipe.host_threads)) if pipe.raw_images > 0 else
                              ("host cores round each image chunk to bf16 (azn_host_f32_to_bf16, %d threads) while the previous chunk uploads; "
                               "bf16 NCHW -> NHWC on the device" % pipe.host_threads)) if pipe.narrow else "f32 batch uploaded as it is; f32 NCHW -> bf16 NHWC on the device",
                    "raw_f32_images_per_batch": pipe.raw_images if pipe.narrow else BATCH,
                    "route_timing": pipe.narrow_timing,
                    "f32_upload": None if ms_e2e_f32 is None else {
                        "value": world * BATCH * args.steps / (ms_e2e_f32 / 1e3), "unit": "images/s", "h2d_bytes_per_step": pipe_f32.h2d_bytes,
                        "ms_per_step": ms_e2e_f32 / args.steps, "note": "same call with host_narrow off: the whole f32 batch crosses PCIe"},
                    "bf16_nhwc_input": {"value": world * BATCH * args.steps / (ms_e2e_bf16 / 1e3), "unit": "images/s",
                                        "h2d_bytes_per_step": pipe_bf16.h2d_bytes, "ms_per_step": ms_e2e_bf16 / args.steps,
                                        "note": "same call, host maps already bf16 NHWC (the engine's storage format): half the PCIe bytes, no conversion kernel"}},
            "gpu_launches": launches,
            "roofline": {"kernel": "fc_gemm_kernel<256,2> int6 25088->4096, all %d launches of a step (time-weighted)" % len(int6), "bound": "tensor",
                         "achieved": int6_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": int6_tf / tf_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": which + " (sustained bf16)",
                         "ms_per_step_in_kernel": int6_ms, "share_of_step": int6_ms / (ms_total / args.steps),
                         "deepest_launch": {"m_rows": top["m"], "ms": top["ms"], "achieved": top["tflops"], "frac": top["tflops"] / tf_peak},
                         "whole_step": {"flops": step_flops, "achieved": step_flops / (ms_total / args.steps * 1e-3) / 1e12,
                                        "frac": step_flops / (ms_total / args.steps * 1e-3) / 1e12 / tf_peak},
                         "per_level": prof["levels"], "hbm_peak_gbs": hbm_peak},
        }
        if args.job:
            line["run"]["job"] = "%d batches of 64 per rank; lists -> rank 0: %s" % (
                args.steps, "peer window + closing fence" if gather_route == "peer" else "one NCCL all_gather at the end")
        if world > 1:
            line["gather_ms"] = {"route": gather_route, "resident": gather_resident_ms, "e2e": gather_ms[0],
                                 "bytes_per_rank": int(sum(t.numel() * t.element_size() for c in collectors for t in (c.boxes, c.scores, c.counts)))}
        threads = os.cpu_count() or 1
        runner = None
        if not args.no_cpu_baseline:
            runner = OracleRunner(weights, threads)
            runner.run(1)
            runner.regions = runner.images = 0
            t0 = time.perf_counter()
            n_cpu = 0
            while n_cpu < 1024 and time.perf_counter() - t0 < 12.0:        # a bounded sample: >= 12 s of host work
                runner.run(16)
                n_cpu += 16
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_cpu / dt, "unit": "images/s", "cores": threads, "kind": runner.kind,
                                    "sample": "%d images of the same workload in %.1f s (%s); %.0f regions/image" % (
                                        n_cpu, dt, runner.what, runner.regions / n_cpu)}
        if not args.no_extra and world == 1:
            try:
                line["parity"] = parity_block(eng, dev_sets[0], host_sets[0], weights)
            except Exception as e:
                line["parity"] = {"error": repr(e)}
            try:
                line["extra"], line["e2e_entry"] = extra_blocks(dev, head, hbm_peak)
            except Exception as e:
                line["extra"] = {"error": repr(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
