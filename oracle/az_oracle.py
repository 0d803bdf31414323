"""oracle/az_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + the plain-C helpers in oracle/caffe_layers.c) of the
reference's adaptive-search hot path, written from the reference's behaviour and
cited function by function.  All citations are relative to /root/reference.

Only tests/, __graft_entry__.smoke() and bench.py's `cpu_baseline` /
`--impl reference` legs may import this module.  The product package
(aznet_b200/) must never import it: a product path that routes through the
oracle voids every parity claim.

Pinning: the reference has NO tests for this path, so
the pin is the reference's own code executed in the authoring container:
  * divide_region/_sift_dup and nms are checked against lib/utils/div.pyx and
    lib/utils/nms.pyx compiled unmodified from /root/reference (oracle/build_ref.py
    -> oracle/_ref/*.so) and against tests/golden/{div,nms}_*.npz made by them;
  * _bbox_pred/_clip_boxes/_unwrap_adj_pred/_az_forward/im_propose/test_net
    selection are checked against lib/detect/test.py (mechanically converted
    py2->py3 into oracle/_ref/, never committed) through tests/golden/search_*.npz;
  * frcnn_forward / test_net_select / apply_nms are checked against the reference's own
    im_detect, test_net (detections.pkl) and apply_nms run by oracle/gen_golden.py --only
    detect with HashDetNet (tests/golden/detect.npz; bit for bit);
  * the Caffe layers: Caffe as a whole cannot be built here (no glog/gflags/boost/BLAS/protoc),
    but the layer SOURCES on the path compile unmodified against the stand-in framework
    headers of oracle/caffe_shim (oracle/build_ref.py -> oracle/_ref/libcaffe_layers_ref.so):
    ROIPooling (+argmax), GRN, Sigmoid, ReLU, Softmax, MAX Pooling are pinned BIT FOR BIT to
    Forward_cpu of those sources (tests/golden/caffe_layers.npz + fresh random inputs);
    InnerProduct's layer code is pinned the same way, its fp32 summation order is not (the
    reference's BLAS is un-vendored: ATLAS | MKL | OpenBLAS, no pinned version), so the fc
    layers stay tolerance-based, as north_star states.
"""
from __future__ import annotations

import ctypes
import heapq
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/caffe_layers.c -> oracle/liboracle.so (gcc, no deps)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "caffe_layers.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        c = ctypes
        L.azo_roi_pool_fwd.restype = c.c_int
        L.azo_roi_pool_fwd.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int,
                                       c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_float,
                                       c.c_void_p, c.c_void_p]
        L.azo_sigmoid.restype = None
        L.azo_sigmoid.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t]
        L.azo_softmax.restype = None
        L.azo_softmax.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t, c.c_int]
        L.azo_nms.restype = c.c_int64
        L.azo_nms.argtypes = [c.c_void_p, c.c_int64, c.c_int64, c.c_void_p, c.c_double, c.c_void_p]
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------------------
# configuration: the cfg keys the hot path reads (lib/detect/config.py:100-216)
# --------------------------------------------------------------------------------------
@dataclass
class OracleCfg:
    TEST_SCALES: tuple = (600,)          # config.py:109
    TEST_MAX_SIZE: int = 1000            # config.py:112  (voc.yml/coco.yml: 800)
    TEST_NMS: float = 0.5                # config.py:116
    NUM_PROPOSALS: int = 300             # config.py:129 via cfg_set_mode('Test') :272-280
    Tz: float = 0.5                      # injected by cfg_set_mode; no default in the reference
    Tc: float = 0.05                     # config.py:166
    FIXED_PROPOSAL_NUM: bool = True      # config.py:167
    APPEND_BOXES: bool = False           # config.py:176
    MIN_SIDE: object = 10                # config.py:183 (an int: Python-2 `/` is floor division)
    BATCH_SIZE: int = 10000              # config.py:186 (voc.yml: 1000)
    DEDUP_BOXES: float = 1. / 16.        # config.py:203
    EPS: float = 1e-14                   # config.py:213
    SPATIAL_SCALE: float = 0.0625        # models/Pascal/VGG16/az-net/test_fc.prototxt:23
    POOLED: int = 7                      # test_fc.prototxt:21-22
    extra: dict = field(default_factory=dict)


# --------------------------------------------------------------------------------------
# Caffe layers
# --------------------------------------------------------------------------------------
def roi_pool_fwd(feat, rois, pooled=7, spatial_scale=0.0625, want_argmax=False):
    """caffe-fast-rcnn/src/caffe/layers/roi_pooling_layer.cpp:46-125 (see caffe_layers.c)."""
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    n, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.empty((R, C, pooled, pooled), dtype=np.float32)
    amax = np.empty((R, C, pooled, pooled), dtype=np.int32) if want_argmax else None
    rc = _lib().azo_roi_pool_fwd(feat.ctypes.data, n, C, H, W, rois.ctypes.data, R, pooled, pooled,
                                 ctypes.c_float(spatial_scale), out.ctypes.data,
                                 amax.ctypes.data if want_argmax else None)
    if rc != 0:
        raise RuntimeError("roi batch index out of range (Caffe CHECK would abort)")
    return (out, amax) if want_argmax else out


def sigmoid(x):
    """caffe-fast-rcnn/src/caffe/layers/sigmoid_layer.cpp:11-13."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    _lib().azo_sigmoid(x.ctypes.data, y.ctypes.data, x.size)
    return y


def softmax(x):
    """caffe-fast-rcnn/src/caffe/layers/softmax_layer.cpp:28-60, rows of [R, C]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    _lib().azo_softmax(x.ctypes.data, y.ctypes.data, x.shape[0], x.shape[1])
    return y


def inner_product(x, w, b, threads=None):
    """caffe-fast-rcnn/src/caffe/layers/inner_product_layer.cpp:80-93:
    top = bottom . W^T (cblas_sgemm, util/math_functions.cpp:13-21) + 1 . b^T, fp32.
    The BLAS is un-vendored in the reference (ATLAS|MKL|OpenBLAS, no pinned version);
    torch-CPU sgemm (MKL) is the same class of library."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    wt = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
    bt = torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32))
    return torch.addmm(bt, xt, wt.t()).numpy()


def relu(x):
    """caffe-fast-rcnn/src/caffe/layers/relu_layer.cpp:16-19."""
    return np.maximum(x, np.float32(0))


def max_pool_ceil(x, kernel=2, stride=2):
    """PoolingLayer (MAX, pad 0), caffe-fast-rcnn/src/caffe/layers/pooling_layer.cpp: output size
    ceil((H - kernel) / stride) + 1 (:93-96), windows clipped to the map (:152-155), strict > from -FLT_MAX (:157-166;
    a NaN never wins).  x f32 [n,C,H,W]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, C, H, W = x.shape
    ph = int(np.ceil(np.float32(H - kernel) / stride)) + 1
    pw = int(np.ceil(np.float32(W - kernel) / stride)) + 1
    out = np.full((n, C, ph, pw), -np.finfo(np.float32).max, np.float32)
    for dy in range(kernel):
        for dx in range(kernel):
            v = x[:, :, dy::stride, dx::stride][:, :, :ph, :pw]
            o = out[:, :, :v.shape[2], :v.shape[3]]
            np.copyto(o, v, where=v > o)
    return out


def grn(x):
    """GRNLayer::Forward_cpu, caffe-fast-rcnn/src/caffe/layers/grn_layer.cpp:27-56, on a blob [N, C, H, W]:
    square (caffe_sqr), sum across channels (caffe_cpu_gemv with a ones vector; fp32, BLAS summation order),
    root (caffe_powx 0.5), divide every channel by the per-position norm (caffe_div).  No epsilon: an all-zero
    position yields 0/0 = NaN, as in the reference."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    sq = x * x
    norm = np.power(sq.sum(axis=1, dtype=np.float32, keepdims=True), np.float32(0.5)).astype(np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (x / norm).astype(np.float32)


def skip_pool5(w, maps, rois, pooled=7, scales=(0.25, 0.125, 0.0625), names=("conv3_3", "conv4_3", "conv5_3"), act_round=None,
               threads=None):
    """models/COCO/VGG16_skip/frcnn/test_fc.prototxt:28-142: roi_pool{3,4,5} -> GRN -> concat (axis 1) -> Power
    (scale 1000) -> conv_pool5 (1x1 convolution = per-position inner product over channels) -> ReLU.
    Returns pool5 [R, c_out, 7, 7] float32.  act_round (e.g. round_bf16) emulates the product's bf16 storage of the
    concat operand."""
    rnd = act_round or (lambda v: v)
    cat = np.concatenate([grn(roi_pool_fwd(maps[n], rois, pooled, sc)) for n, sc in zip(names, scales)], axis=1)
    cat = rnd((cat * np.float32(1000.0)).astype(np.float32))                     # power_layer.cpp: y = scale * x (power 1)
    R, ctot = cat.shape[0], cat.shape[1]
    wc, bc = w["conv_pool5"]
    x = cat.transpose(0, 2, 3, 1).reshape(R * pooled * pooled, ctot)            # one row per pooled position
    y = relu(inner_product(x, wc.reshape(wc.shape[0], ctot), bc, threads))
    return np.ascontiguousarray(y.reshape(R, pooled, pooled, -1).transpose(0, 3, 1, 2))


def round_bf16(x):
    """float32 -> nearest-even bfloat16 -> float32 (the storage format of the product's activations)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u + np.uint32(0x7fff) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xffff0000)).astype(np.uint32)
    r = np.where((u & np.uint32(0x7f800000)) == np.uint32(0x7f800000), u, r)        # inf / nan pass through
    return r.view(np.float32).reshape(np.shape(x))


class _Blob:
    def __init__(self):
        self.shape = ()

    def reshape(self, *shape):
        self.shape = tuple(shape)

    @property
    def num(self):
        return self.shape[0] if self.shape else 0


class _PortLayers:
    roi_pool_fwd = staticmethod(lambda *a, **k: roi_pool_fwd(*a, **k))
    relu = staticmethod(lambda x: relu(x))
    sigmoid = staticmethod(lambda x: sigmoid(x))
    softmax = staticmethod(lambda x: softmax(x))


class _RefLayers:
    def __init__(self):
        from . import ref_caffe
        if not ref_caffe.available():
            raise RuntimeError("oracle/_ref/libcaffe_layers_ref.so is not built (oracle/build_ref.py)")
        self.roi_pool_fwd, self.relu, self.sigmoid, self.softmax = ref_caffe.roi_pool_fwd, ref_caffe.relu, ref_caffe.sigmoid, ref_caffe.softmax


class OracleNet:
    """Duck-typed caffe.Net (caffe-fast-rcnn/python/caffe/pycaffe.py:52-95) running the
    fc part of AZ-Net (models/Pascal/VGG16/az-net/test_fc.prototxt:14-232) or of the
    Fast R-CNN detector (models/*/VGG16/frcnn/test_fc.prototxt:14-145) on the CPU.

    weights: dict name -> (W [N,K] f32, b [N] f32) with the prototxt layer names.
    kind:    'az' | 'frcnn'.  A `backbone` callable (data blob -> conv5_3) turns it into the
             'full' net; without one it is the 'fc' net whose inputs are (conv5_3, rois).
    """

    def __init__(self, weights, kind="az", backbone=None, name="oracle", cfg=None, threads=None, act_round=None, layers="port"):
        # layers="ref": ROIPooling / ReLU / Sigmoid / Softmax run the reference's OWN layer sources compiled into
        # oracle/_ref/libcaffe_layers_ref.so (oracle/ref_caffe.py); InnerProduct stays the threaded sgemm below (the
        # reference's BLAS is un-vendored).  Used by bench.py's reference arm.
        self.L = _RefLayers() if layers == "ref" else _PortLayers()
        self.w = weights
        # optional storage rounding of the hidden activations (e.g. round_bf16): lets a test separate the
        # product's bf16 activation storage from everything else.  None = the reference's fp32 blobs.
        self.act_round = act_round or (lambda v: v)
        self.kind = kind
        self.backbone = backbone
        self.name = name
        self.cfg = cfg or OracleCfg()
        self.threads = threads
        self.conv_names = ["conv3_3", "conv4_3", "conv5_3"] if kind == "frcnn_skip" else ["conv5_3"]
        self.inputs = ["data", "rois"] if backbone is not None else self.conv_names + ["rois"]
        self.outputs = ["zoom_prob", "adj_prob", "adj_bbox"] if kind == "az" else ["cls_prob", "bbox_pred"]
        self.blobs = {k: _Blob() for k in self.inputs + self.conv_names}
        self.stats = {"pool_s": 0.0, "fc_s": 0.0}

    def forward(self, blobs=None, **kwargs):
        import time
        if set(kwargs.keys()) != set(self.inputs):                       # pycaffe.py:83-84
            raise Exception("Input blob arguments do not match net inputs.")
        for k, v in kwargs.items():                                      # pycaffe.py:88-89
            if v.shape[0] != self.blobs[k].num:
                raise Exception("Input is not batch sized")
        rois = kwargs["rois"]
        t0 = time.perf_counter()
        if self.kind == "frcnn_skip":
            # the skip-layer detector: `backbone` returns a dict of the three maps (test.prototxt of VGG16_skip)
            conv = self.backbone(kwargs["data"]) if self.backbone is not None else {n: kwargs[n] for n in self.conv_names}
            pool5 = skip_pool5(self.w, conv, rois, self.cfg.POOLED, act_round=self.act_round, threads=self.threads)
        else:
            conv = self.backbone(kwargs["data"]) if self.backbone is not None else kwargs["conv5_3"]
            pool5 = self.L.roi_pool_fwd(conv, rois, self.cfg.POOLED, self.cfg.SPATIAL_SCALE)
        t1 = time.perf_counter()
        x = pool5.reshape(pool5.shape[0], -1)            # K index = c*49 + ph*7 + pw (Q12)
        ip = lambda name, v: inner_product(v, self.w[name][0], self.w[name][1], self.threads)
        out = {}
        rnd = self.act_round
        relu, sigmoid, softmax = self.L.relu, self.L.sigmoid, self.L.softmax
        if self.kind == "az":
            h6 = rnd(relu(ip("int6", x)))                 # dropout in TEST phase = identity
            h71 = rnd(relu(ip("int7_1", h6)))
            h72 = rnd(relu(ip("int7_2", h6)))
            out["adj_prob"] = sigmoid(ip("adj_score", h71))
            out["adj_bbox"] = ip("adj_bbox", h71)
            out["zoom_prob"] = sigmoid(ip("zoom_score", h72))
        else:
            h6 = rnd(relu(ip("fc6", x)))
            h7 = rnd(relu(ip("fc7", h6)))
            out["cls_prob"] = softmax(ip("cls_score", h7))
            out["bbox_pred"] = ip("bbox_pred", h7)
        self.stats["pool_s"] += t1 - t0
        self.stats["fc_s"] += time.perf_counter() - t1
        for b in (blobs or []):
            if self.kind == "frcnn_skip" and b in self.conv_names:
                out[b] = conv[b]
            elif b == "conv5_3":
                out[b] = conv
            elif b == "pool5":
                out[b] = pool5
        return out


# --------------------------------------------------------------------------------------
# lib/utils/div.pyx
# --------------------------------------------------------------------------------------
def sift_dup(regions, min_height):
    """lib/utils/div.pyx:78-88."""
    regions = np.asarray(regions, dtype=np.float64).reshape(-1, 4)
    v = np.array([1, 1e3, 1e6, 1e9], dtype=np.float64)
    hashes = np.round(regions / min_height).dot(v)
    _, index = np.unique(hashes, return_index=True)
    return regions[index, :]


def divide_region(regions, min_height=10.0):
    """lib/utils/div.pyx:15-76: split every region into a 2 x num_long grid of
    near-square cells plus the 1 x (num_long-1) cells offset by half a cell, then
    _sift_dup.  Arithmetic is float64, in the operation order of :47-72."""
    regions = np.asarray(regions, dtype=np.float64).reshape(-1, 4)
    out = []
    for x1, y1, x2, y2 in regions:
        lengths = (x2 - x1 + 1.0, y2 - y1 + 1.0)                      # :32-33
        min_ind = 0 if lengths[0] <= lengths[1] else 1                # :35 argmin, tie -> 0
        max_ind = 1 - min_ind
        l_short = lengths[min_ind] / 2                                # :41
        num_long = int(lengths[max_ind] / l_short)                    # :43 (truncation)
        l_long = lengths[max_ind] / num_long                          # :44
        sub = np.zeros((2 * num_long + (num_long - 1), 4))            # :46-47
        for k in range(2):
            for j in range(num_long):
                if min_ind == 0:                                      # width is the short side
                    sub[k * num_long + j] = (k * l_short, j * l_long, (k + 1) * l_short, (j + 1) * l_long)
                else:
                    sub[k * num_long + j] = (j * l_long, k * l_short, (j + 1) * l_long, (k + 1) * l_short)
        offset = 2 * num_long
        h_short = l_short / 2
        h_long = l_long / 2
        for j in range(num_long - 1):                                 # k == 0 only (num_short-1 == 1)
            if min_ind == 0:
                sub[j + offset] = (h_short, j * l_long + h_long, l_short + h_short, (j + 1) * l_long + h_long)
            else:
                sub[j + offset] = (j * l_long + h_long, h_short, (j + 1) * l_long + h_long, l_short + h_short)
        # the reference forms k*l_short + h_short with k == 0, i.e. 0*l_short + h_short == h_short
        # and (k+1)*l_short + h_short == 1*l_short + h_short == l_short + h_short exactly.
        sub[:, [0, 2]] += x1                                          # :71-72
        sub[:, [1, 3]] += y1
        out.append(sub)
    regions_out = np.vstack(out) if out else np.zeros((0, 4))
    return sift_dup(regions_out, min_height)


# --------------------------------------------------------------------------------------
# lib/utils/nms.pyx
# --------------------------------------------------------------------------------------
def nms(dets, thresh, stable_ties=True):
    """lib/utils/nms.pyx:17-68.  `order = scores.argsort()[::-1]` (:25); with
    stable_ties the ascending sort is the stable one, which is what the reference's
    default introsort yields whenever scores are unique (benchmarks and goldens use
    tie-free scores)."""
    if not (isinstance(dets, np.ndarray) and dets.dtype == np.float32 and dets.ndim == 2):
        raise ValueError("Buffer dtype mismatch, expected 'float32_t'")   # Cython typed buffer, :17
    if not isinstance(thresh, float):
        raise TypeError("Argument 'thresh' has incorrect type (expected float)")
    n = dets.shape[0]
    d = np.ascontiguousarray(dets)
    scores = d[:, 4]
    order = scores.argsort(kind="stable" if stable_ties else None)[::-1].astype(np.int64)
    order = np.ascontiguousarray(order)
    keep = np.empty(max(n, 1), dtype=np.int64)
    nk = _lib().azo_nms(d.ctypes.data, d.shape[1], n, order.ctypes.data, float(thresh), keep.ctypes.data)
    return [int(i) for i in keep[:nk]]


def apply_nms(all_boxes, thresh):
    """lib/detect/test.py:467-484."""
    num_classes = len(all_boxes)
    num_images = len(all_boxes[0])
    nms_boxes = [[[] for _ in range(num_images)] for _ in range(num_classes)]
    for c in range(num_classes):
        for i in range(num_images):
            dets = all_boxes[c][i]
            if isinstance(dets, list) and dets == []:
                continue
            keep = nms(dets, thresh)
            if len(keep) == 0:
                continue
            nms_boxes[c][i] = dets[keep, :].copy()
    return nms_boxes


# --------------------------------------------------------------------------------------
# lib/detect/test.py
# --------------------------------------------------------------------------------------
def im_scale_for(im_shape, cfg):
    """Scale logic of _get_image_blob, lib/detect/test.py:40-52 (one scale per TEST.SCALES)."""
    size_min = min(im_shape[0], im_shape[1])
    size_max = max(im_shape[0], im_shape[1])
    scales = []
    for target in cfg.TEST_SCALES:
        s = float(target) / float(size_min)
        if np.round(s * size_max) > cfg.TEST_MAX_SIZE:
            s = float(cfg.TEST_MAX_SIZE) / float(size_max)
        scales.append(s)
    return np.array(scales)


def get_rois_blob(im_rois, scales):
    """_get_rois_blob + _project_im_rois, lib/detect/test.py:61-97 (single-scale branch :92-95;
    the multi-scale branch :83-91 is restated too)."""
    im_rois = np.asarray(im_rois).astype(np.float64, copy=False)
    if len(scales) > 1:
        widths = im_rois[:, 2] - im_rois[:, 0] + 1
        heights = im_rois[:, 3] - im_rois[:, 1] + 1
        areas = widths * heights
        scaled = areas[:, None] * (scales[None, :] ** 2)
        levels = np.abs(scaled - 224 * 224).argmin(axis=1)[:, None]
    else:
        levels = np.zeros((im_rois.shape[0], 1), dtype=np.int64)
    rois = im_rois * scales[levels]
    return np.hstack((levels, rois)).astype(np.float32, copy=False)


def bbox_pred(boxes, box_deltas, eps=1e-14):
    """_bbox_pred, lib/detect/test.py:106-139.  float64 except np.exp on the float32 deltas."""
    if boxes.shape[0] == 0:
        return np.zeros((0, box_deltas.shape[1]))
    boxes = boxes.astype(np.float64, copy=False)
    widths = boxes[:, 2] - boxes[:, 0] + eps
    heights = boxes[:, 3] - boxes[:, 1] + eps
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    dx, dy, dw, dh = (box_deltas[:, i::4] for i in range(4))
    pcx = dx * widths[:, None] + ctr_x[:, None]
    pcy = dy * heights[:, None] + ctr_y[:, None]
    pw = np.exp(dw) * widths[:, None]
    ph = np.exp(dh) * heights[:, None]
    pred = np.zeros(box_deltas.shape)
    pred[:, 0::4] = pcx - 0.5 * pw
    pred[:, 1::4] = pcy - 0.5 * ph
    pred[:, 2::4] = pcx + 0.5 * pw
    pred[:, 3::4] = pcy + 0.5 * ph
    return pred


def clip_boxes(boxes, im_shape):
    """_clip_boxes, lib/detect/test.py:141-151 (in place)."""
    boxes[:, 0::4] = np.maximum(boxes[:, 0::4], 0)
    boxes[:, 1::4] = np.maximum(boxes[:, 1::4], 0)
    boxes[:, 2::4] = np.minimum(boxes[:, 2::4], im_shape[1] - 1)
    boxes[:, 3::4] = np.minimum(boxes[:, 3::4], im_shape[0] - 1)
    return boxes


def unwrap_adj_pred(boxes, scores, min_side):
    """_unwrap_adj_pred, lib/detect/test.py:171-187."""
    scores = scores.ravel()
    b = np.vstack((boxes[:, 0::4].ravel(), boxes[:, 1::4].ravel(),
                   boxes[:, 2::4].ravel(), boxes[:, 3::4].ravel())).transpose()
    heights = b[:, 3] - b[:, 1] + 1
    widths = b[:, 2] - b[:, 0] + 1
    keep = np.where(np.minimum(heights, widths) >= min_side)[0]
    return b[keep, :], scores[keep]


def _dedup(rois_blob, dedup):
    """lib/detect/test.py:212-218."""
    v = np.array([1, 1e3, 1e6, 1e9, 1e12])
    hashes = np.round(rois_blob * dedup).dot(v)
    _, index, inv = np.unique(hashes, return_index=True, return_inverse=True)
    return index, inv


def az_forward(net, im_shape, all_boxes, conv, cfg, data_blob=None):
    """_az_forward, lib/detect/test.py:189-257.  `im_shape` replaces `im` (only im.shape and
    the image blob are used); `data_blob` is what _get_image_blob would produce for the 'full'
    net, and is not needed once `conv` is cached."""
    bs = cfg.BATCH_SIZE
    nb = int(np.ceil(all_boxes.shape[0] / float(bs)))
    z_all, a_all, c_all = np.zeros((0,)), np.zeros((0, 4)), np.zeros((0,))
    scales = im_scale_for(im_shape, cfg)
    for bid in range(nb):
        boxes = all_boxes[bs * bid:min(all_boxes.shape[0], bs * (bid + 1)), 0:4]
        rois = get_rois_blob(boxes, scales)
        inv = None
        if cfg.DEDUP_BOXES > 0:
            index, inv = _dedup(rois, cfg.DEDUP_BOXES)
            rois = rois[index, :]
            boxes = boxes[index, :]
        if conv is None or "fc" not in net.keys():
            net["full"].blobs["data"].reshape(*data_blob.shape)
            net["full"].blobs["rois"].reshape(*rois.shape)
            out = net["full"].forward(data=data_blob.astype(np.float32, copy=False),
                                      rois=rois.astype(np.float32, copy=False), blobs=["conv5_3"])
            conv = {"conv5_3": out["conv5_3"]}
        else:
            net["fc"].blobs["conv5_3"].reshape(*conv["conv5_3"].shape)
            net["fc"].blobs["rois"].reshape(*rois.shape)
            out = net["fc"].forward(rois=rois.astype(np.float32, copy=False), conv5_3=conv["conv5_3"])
        z = out["zoom_prob"]
        scores = out["adj_prob"]
        pred = clip_boxes(bbox_pred(boxes, out["adj_bbox"], cfg.EPS), im_shape)
        if cfg.DEDUP_BOXES > 0:
            scores = scores[inv, :]
            pred = pred[inv, :]
            z = z[inv].ravel()
        else:
            z = z.ravel()
        a, c = unwrap_adj_pred(pred, scores, cfg.MIN_SIDE)
        z_all = np.hstack((z_all, z))
        a_all = np.vstack((a_all, a))
        c_all = np.hstack((c_all, c))
    return z_all, a_all, c_all, conv


def search_depth(im_shape, cfg):
    """K of lib/detect/test.py:363-368.  `side/cfg.SEAR.MIN_SIDE` is Python-2 division:
    floor division when MIN_SIDE is an int."""
    side = int(min(im_shape[0], im_shape[1]))
    q = side // cfg.MIN_SIDE if isinstance(cfg.MIN_SIDE, (int, np.integer)) else side / cfg.MIN_SIDE
    return int(np.log2(q) + 1.0)


APPEND_TEMP = np.transpose(np.array([[[0, 0, 1, 1], [-0.25, 0, 1, 1], [0, 0, 1.25, 1], [0, -0.25, 1, 1], [0, 0, 1, 1.25],
                                      [-0.125, -0.125, 1.125, 1.125], [0.125, 0.125, 0.875, 0.875]]]), axes=[0, 2, 1])   # config.py:177-183


def append_boxes(boxes, cfg):
    """_append_boxes, lib/detect/test.py:320-344: template boxes around every proposal, template-major, then
    `_sift_dup` on a grid of 1 / DEDUP_BOXES pixels."""
    n, ns = boxes.shape[0], APPEND_TEMP.shape[2]
    w, h = boxes[:, [2]] - boxes[:, [0]], boxes[:, [3]] - boxes[:, [1]]
    Lm = np.hstack((w, h, w, h))[:, :, np.newaxis]
    delta = np.hstack((boxes[:, [0]], boxes[:, [1]], boxes[:, [0]], boxes[:, [1]]))[:, :, np.newaxis]
    subs = np.transpose(Lm * APPEND_TEMP + delta, [2, 0, 1]).reshape((n * ns, 4)).astype(np.float64, copy=False)
    return sift_dup(subs, 1 / cfg.DEDUP_BOXES)


def im_propose(net, im_shape, cfg, conv=None, data_blob=None, num_proposals=None, return_scores=False,
               trace=None):
    """im_propose, lib/detect/test.py:346-414 (with cfg.APPEND_BOXES the appended, clipped boxes of :403-406; the
    returned scores then still belong to the un-appended selection).  Returns Y [n,4] float64 (and the matching
    scores / a per-level trace for the parity tests)."""
    B = np.array([[0, 0, im_shape[1] - 1.0, im_shape[0] - 1.0]])
    Y = np.zeros((0, 4))
    a_scores = np.zeros((0,))
    num_eval = 0
    K = search_depth(im_shape, cfg)
    k = 0
    for k in range(1, K):
        zoom, boxes, c, conv = az_forward(net, im_shape, B, conv, cfg, data_blob)
        num_eval += B.shape[0]
        Y = np.vstack((Y, boxes))
        a_scores = np.hstack((a_scores, c))
        if k == 1:
            zoom[0] = 1.0
        ind_z = np.where(zoom >= cfg.Tz)[0]
        if trace is not None:
            trace.append({"B": B.copy(), "zoom": zoom.copy(), "n_boxes": boxes.shape[0]})
        Z = B[ind_z, :]
        if Z.shape[0] == 0:
            break
        B = divide_region(Z, float(cfg.MIN_SIDE))
    if (not cfg.FIXED_PROPOSAL_NUM) and num_proposals is None:
        ind_a = np.where(a_scores >= cfg.Tc)[0]
    else:
        if num_proposals is None:
            num_proposals = cfg.NUM_PROPOSALS
        ind_a = np.argsort(-a_scores, kind="stable")[:min(num_proposals, Y.shape[0])]
    info = {"num_eval": num_eval, "depth": k, "Y_all": Y, "scores_all": a_scores}
    Y = Y[ind_a, :]
    if cfg.APPEND_BOXES:
        Y = clip_boxes(append_boxes(Y, cfg), im_shape)
    if return_scores:
        return Y, a_scores[ind_a], info
    return Y


def im_propose_tune(net, im_shape, cfg, conv=None, data_blob=None):
    """The diagnostic im_propose of lib/detect/tune.py:256-316: K levels (`for k in xrange(K)`), Tz = 0 for the
    first level and cfg.SEAR.Tz afterwards (:278, :305), no forced root, anchor history Bhis (:298).
    Returns (hstack(Y, scores) [n,5], Bhis [m,5], info)."""
    B = np.array([[0, 0, im_shape[1] - 1.0, im_shape[0] - 1.0]])
    Bhis = np.zeros((0, 5))
    Y = np.zeros((0, 4))
    a_scores = np.zeros((0,))
    num_eval = 0
    K = search_depth(im_shape, cfg)
    Tz = 0
    k = 0
    for k in range(K):
        zoom, boxes, c, conv = az_forward(net, im_shape, B, conv, cfg, data_blob)
        num_eval += B.shape[0]
        Y = np.vstack((Y, boxes))
        a_scores = np.hstack((a_scores, c))
        ind_z = np.where(zoom >= Tz)[0]
        Z = B[ind_z, :]
        Bhis = np.vstack((Bhis, np.hstack((B, zoom[:, np.newaxis]))))
        if Z.shape[0] == 0:
            break
        B = divide_region(Z, float(cfg.MIN_SIDE))
        Tz = cfg.Tz
    ind_a = np.argsort(-a_scores, kind="stable")[:min(cfg.NUM_PROPOSALS, Y.shape[0])]
    info = {"num_eval": num_eval, "depth": k}
    return np.hstack((Y[ind_a, :], a_scores[ind_a, np.newaxis])), Bhis, info


def tune_thresh(histories, anchors_per_img=20):
    """tune_thresh's heap loop (lib/detect/tune.py:318-366) over the per-image anchor histories [m_i, 5]:
    the zoom threshold that keeps num_images * cfg.TRAIN.ANCHORS_PER_IMG anchors."""
    import heapq
    max_per_set = len(histories) * anchors_per_img
    top_scores, thresh = [], -np.inf
    for h in histories:
        scores = h[:, -1]
        for val in scores[np.where(scores > thresh)[0]]:
            heapq.heappush(top_scores, val)
        if len(top_scores) > max_per_set:
            while len(top_scores) > max_per_set:
                heapq.heappop(top_scores)
            thresh = top_scores[0]
    return thresh


def frcnn_forward(net, im_shape, all_boxes, num_classes, conv, cfg, data_blob=None):
    """_frcnn_forward, lib/detect/test.py:259-318."""
    bs = cfg.BATCH_SIZE
    nb = int(np.ceil(all_boxes.shape[0] / float(bs)))
    all_pred = np.zeros((0, 4 * num_classes))
    all_scores = np.zeros((0, num_classes))
    scales = im_scale_for(im_shape, cfg)
    for bid in range(nb):
        boxes = all_boxes[bs * bid:min(all_boxes.shape[0], bs * (bid + 1)), 0:4]
        rois = get_rois_blob(boxes, scales)
        inv = None
        if cfg.DEDUP_BOXES > 0:
            index, inv = _dedup(rois, cfg.DEDUP_BOXES)
            rois = rois[index, :]
            boxes = boxes[index, :]
        if conv is None or "fc" not in net.keys():
            net["full"].blobs["data"].reshape(*data_blob.shape)
            net["full"].blobs["rois"].reshape(*rois.shape)
            out = net["full"].forward(data=data_blob.astype(np.float32, copy=False),
                                      rois=rois.astype(np.float32, copy=False), blobs=["conv5_3"])
            conv = {"conv5_3": out["conv5_3"]}
        else:
            net["fc"].blobs["conv5_3"].reshape(*conv["conv5_3"].shape)
            net["fc"].blobs["rois"].reshape(*rois.shape)
            out = net["fc"].forward(rois=rois.astype(np.float32, copy=False), conv5_3=conv["conv5_3"])
        scores = out["cls_prob"]
        pred = clip_boxes(bbox_pred(boxes, out["bbox_pred"], cfg.EPS), im_shape)
        if cfg.DEDUP_BOXES > 0:
            scores = scores[inv, :]
            pred = pred[inv, :]
        all_scores = np.vstack((all_scores, scores))
        all_pred = np.vstack((all_pred, pred))
    return all_scores, all_pred, conv


def test_net_select(per_image, num_classes, max_per_image=100):
    """Per-class selection of test_net, lib/detect/test.py:549-651.
    per_image: list of (scores [R,C], boxes [R,4C]) or None for images without proposals.
    Returns (all_boxes[cls][img] float32 [n,5] | [], thresh [C])."""
    num_images = len(per_image)
    max_per_set = 800 // (num_classes - 1) * num_images               # :551 Python-2 int division
    thresh = -np.inf * np.ones(num_classes)
    top_scores = [[] for _ in range(num_classes)]
    all_boxes = [[[] for _ in range(num_images)] for _ in range(num_classes)]
    for i, item in enumerate(per_image):
        if item is None:
            continue
        scores, boxes = item
        for j in range(1, num_classes):
            inds = np.where(scores[:, j] > thresh[j])[0]
            cls_scores = scores[inds, j]
            cls_boxes = boxes[inds, j * 4:(j + 1) * 4]
            top = np.argsort(-cls_scores, kind="stable")[:max_per_image]
            cls_scores = cls_scores[top]
            cls_boxes = cls_boxes[top, :]
            for val in cls_scores:
                heapq.heappush(top_scores[j], val)
            if len(top_scores[j]) > max_per_set:
                while len(top_scores[j]) > max_per_set:
                    heapq.heappop(top_scores[j])
                thresh[j] = top_scores[j][0]
            all_boxes[j][i] = np.hstack((cls_boxes, cls_scores[:, None])).astype(np.float32, copy=False)
    for j in range(1, num_classes):
        for i, item in enumerate(per_image):
            if item is None:
                continue
            inds = np.where(all_boxes[j][i][:, -1] > thresh[j])[0]
            all_boxes[j][i] = all_boxes[j][i][inds, :]
    return all_boxes, thresh


# ---- backbone: image blob and VGG16 conv1_1 .. conv5_3 -------------------------------------------
PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])          # lib/detect/config.py:210


def resize_linear_f32(src, fx, fy):
    """cv2.resize(src f32 HxWxC, None, None, fx, fy, INTER_LINEAR) restated in numpy (OpenCV 4.x
    modules/imgproc/src/resize.cpp: coordinate tables in float from a double scale, horizontal pass
    S[x0]*(1-a) + S[x1]*a, vertical pass T0*(1-b) + T1*b, all float32; destination size = cvRound(size * f)).
    Pinned against cv2 itself by tests/test_oracle.py::test_image_blob_matches_cv2 (<= 1 ulp-level differences:
    cv2's SIMD path may contract a multiply-add)."""
    h0, w0 = src.shape[:2]
    hs, ws = int(np.rint(h0 * fy)), int(np.rint(w0 * fx))
    sx_scale, sy_scale = 1.0 / fx, 1.0 / fy

    def table(n_dst, n_src, scale, clamp_frac):
        d = np.arange(n_dst, dtype=np.float64)
        f = (d + 0.5) * scale - 0.5                      # source coordinate, double
        i = np.floor(f).astype(np.int64)
        a = (f - i).astype(np.float32)                   # fraction taken in double, THEN rounded to float
        if clamp_frac:                                   # x: the fraction is zeroed where the index is clamped
            a[i < 0] = 0
            i[i < 0] = 0
            a[i >= n_src - 1] = 0
            i[i >= n_src - 1] = n_src - 1
            return i, np.minimum(i + 1, n_src - 1), a
        return np.clip(i, 0, n_src - 1), np.clip(i + 1, 0, n_src - 1), a      # y: rows clamped, fraction kept

    x0, x1, ax = table(ws, w0, sx_scale, True)
    y0, y1, ay = table(hs, h0, sy_scale, False)
    src = src.astype(np.float32, copy=False)
    a1 = ax[None, :, None]
    a0 = (np.float32(1) - ax)[None, :, None]
    rows = src[:, x0] * a0 + src[:, x1] * a1             # horizontal pass on every source row (float32)
    b1 = ay[:, None, None]
    b0 = (np.float32(1) - ay)[:, None, None]
    return (rows[y0] * b0 + rows[y1] * b1).astype(np.float32)


def get_image_blob(im, cfg):
    """_get_image_blob + im_list_to_blob (lib/detect/test.py:27-59, lib/utils/blob.py:13-29) for one scale:
    uint8 HxWx3 BGR -> f32 blob [1, 3, Hs, Ws], im_scale."""
    im_orig = im.astype(np.float32, copy=True)
    im_orig -= PIXEL_MEANS
    s = im_scale_for(im.shape, cfg)[0]
    out = resize_linear_f32(im_orig, s, s)
    return np.ascontiguousarray(out.transpose(2, 0, 1)[None]), s


def vgg16_conv5(weights, data, threads=None):
    """conv1_1 .. conv5_3 of models/Pascal/VGG16/az-net/test.prototxt:16-384 in fp32 on the CPU:
    ConvolutionLayer (im2col + sgemm + bias, conv_layer.cpp / base_conv_layer.cpp) = a 3x3 pad-1 correlation,
    in-place ReLU, MAX pooling 2x2/2 in ceil mode (pooling_layer.cpp:81-95).  data f32 [n,3,H,W] -> f32 [n,C,h,w]."""
    import torch
    import torch.nn.functional as F
    if threads:
        torch.set_num_threads(threads)
    x = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
    stages = [2, 2, 3, 3, 3]
    with torch.no_grad():
        for s, n in enumerate(stages, 1):
            for i in range(1, n + 1):
                W, b = weights["conv%d_%d" % (s, i)]
                x = F.relu(F.conv2d(x, torch.from_numpy(np.ascontiguousarray(W)), torch.from_numpy(np.ascontiguousarray(b)), padding=1))
            if s < 5:
                x = F.max_pool2d(x, 2, 2, ceil_mode=True)
    return x.numpy()
