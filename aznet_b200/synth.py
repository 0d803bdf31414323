"""Seeded synthetic inputs for the AZ-Net hot path.

The reference ships no weights, images or thresholds that can be used offline
(tools/prop_az.py:74-92 needs a .caffemodel and a thresh.pkl; the prototxt fillers give
all-zero conv5_3 and tie-only scores), so every benchmark and parity
test runs on arrays made here and handed, as the same bits, to both the CUDA path and
the oracle.  Nothing in this module touches the GPU.
"""
from __future__ import annotations

import numpy as np

# Layer dimensions of models/Pascal/VGG16/az-net/test_fc.prototxt:14-232 and
# models/{Pascal,COCO}/VGG16/frcnn/test_fc.prototxt:14-145.
AZ_DIMS = dict(C=512, pooled=7, h6=4096, h71=1024, h72=256, nsub=11)
FRCNN_DIMS = dict(C=512, pooled=7, h6=4096, h7=4096)


def _normal(rng, shape, std):
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


def make_az_weights(seed=3, C=512, pooled=7, h6=4096, h71=1024, h72=256, nsub=11, zoom_bias=-2.0):
    """He-normal fc weights for the AZ-Net head; dict layer-name -> (W [out,in] f32, b [out] f32).
    Row-major [N, K] like Caffe's InnerProduct blobs (inner_product_layer.cpp:34-37); the K index of
    int6 is c*pooled*pooled + ph*pooled + pw (roi_pooling_layer.cpp:39-40)."""
    rng = np.random.default_rng(seed)
    k6 = C * pooled * pooled
    w = {}
    w["int6"] = (_normal(rng, (h6, k6), np.sqrt(2.0 / k6)), _normal(rng, (h6,), 0.01))
    w["int7_1"] = (_normal(rng, (h71, h6), np.sqrt(2.0 / h6)), _normal(rng, (h71,), 0.01))
    w["int7_2"] = (_normal(rng, (h72, h6), np.sqrt(2.0 / h6)), _normal(rng, (h72,), 0.01))
    w["adj_score"] = (_normal(rng, (nsub, h71), 2.0 * np.sqrt(2.0 / h71)), _normal(rng, (nsub,), 0.01))
    w["adj_bbox"] = (_normal(rng, (4 * nsub, h71), 0.1 * np.sqrt(2.0 / h71)), _normal(rng, (4 * nsub,), 0.01))
    w["zoom_score"] = (_normal(rng, (1, h72), 2.0 * np.sqrt(2.0 / h72)),
                       np.full((1,), zoom_bias, dtype=np.float32))
    return w


def make_frcnn_weights(seed=4, num_classes=21, C=512, pooled=7, h6=4096, h7=4096):
    """He-normal fc weights for the Fast R-CNN head (fc6, fc7, cls_score, bbox_pred)."""
    rng = np.random.default_rng(seed)
    k6 = C * pooled * pooled
    w = {}
    w["fc6"] = (_normal(rng, (h6, k6), np.sqrt(2.0 / k6)), _normal(rng, (h6,), 0.01))
    w["fc7"] = (_normal(rng, (h7, h6), np.sqrt(2.0 / h6)), _normal(rng, (h7,), 0.01))
    w["cls_score"] = (_normal(rng, (num_classes, h7), 2.0 * np.sqrt(2.0 / h7)),
                      _normal(rng, (num_classes,), 0.01))
    w["bbox_pred"] = (_normal(rng, (4 * num_classes, h7), 0.1 * np.sqrt(2.0 / h7)),
                      _normal(rng, (4 * num_classes,), 0.01))
    return w


def make_frcnn_skip_weights(seed=4, num_classes=21, channels=(256, 512, 512), c_out=512, pooled=7, h6=4096, h7=4096):
    """The skip-layer detector head (models/COCO/VGG16_skip/frcnn/test_fc.prototxt): the Fast R-CNN fc weights over a
    c_out-channel pool5 plus conv_pool5, the 1x1 convolution over the GRN-normalised, x1000-scaled concat of the three
    ROI pools -- gaussian std 0.001 like its weight_filler (test_fc.prototxt:124-131), which gives O(1) activations."""
    w = make_frcnn_weights(seed, num_classes, C=c_out, pooled=pooled, h6=h6, h7=h7)
    rng = np.random.default_rng(seed + 1000)
    ctot = int(sum(channels))
    w["conv_pool5"] = (_normal(rng, (c_out, ctot, 1, 1), 0.001), np.zeros((c_out,), np.float32))
    w["skip_channels"] = tuple(int(c) for c in channels)
    return w


def conv_shape(im_h, im_w, im_scale=1.0):
    """conv5_3 spatial size of VGG16 for a scaled image: four ceil-mode 2x2/2 max-pools
    (caffe-fast-rcnn/src/caffe/layers/pooling_layer.cpp:93-95), 3x3 pad-1 convs keep size."""
    h = int(np.round(im_h * im_scale))
    w = int(np.round(im_w * im_scale))
    for _ in range(4):
        h = (h + 1) // 2
        w = (w + 1) // 2
    return h, w


def make_conv_maps(n, C, H, W, seed=7):
    """Post-ReLU conv5_3 stand-ins, f32 NCHW [n, C, H, W]: relu(randn), one seed per image."""
    out = np.empty((n, C, H, W), dtype=np.float32)
    for i in range(n):
        rng = np.random.default_rng(seed + i)
        out[i] = np.maximum(rng.standard_normal((C, H, W), dtype=np.float32), 0)
    return out


def make_images(n, H=600, W=1000, seed=1000):
    """uint8 HWC BGR images (content is irrelevant to throughput)."""
    return [np.random.default_rng(seed + i).integers(0, 256, (H, W, 3), dtype=np.uint8) for i in range(n)]


def make_boxes(n, im_h=600, im_w=1000, seed=3, lo=16.0, hi=400.0):
    """Box generator of SURVEY 8d: centres uniform, w = exp(U(ln lo, ln hi)), h = w*exp(U(-.7,.7)),
    clipped to the image.  float64 [n,4] (x1,y1,x2,y2)."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, im_w - 1, n)
    cy = rng.uniform(0, im_h - 1, n)
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    h = w * np.exp(rng.uniform(-0.7, 0.7, n))
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, im_w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, im_h - 1)
    return b


def make_detect_proposals(counts, im_h, im_w, seed=21, cap=None):
    """Proposal lists for the detection step: f64 [n_img, cap, 4] (zero padded) with, per image, pairs that share a
    feature-space cell but not their image-space box and exact duplicates (the dedup / un-dedup of _frcnn_forward,
    lib/detect/test.py:280-288, 311-313).  Shared by oracle/gen_golden.py --only detect and the tests."""
    cap = max(counts) if cap is None else cap
    boxes = np.zeros((len(counts), max(cap, 1), 4))
    for i, c in enumerate(counts):
        b = make_boxes(c, im_h, im_w, seed=seed + i, lo=12, hi=500)
        if c > 40:
            b[30:40] = b[5:15] + 0.25
            b[c - 3:] = b[:3]
        boxes[i, :c] = b
    return boxes


def make_rois(n, im_h=600, im_w=1000, seed=3, n_img=1):
    """f32 [n,5] ROI blob (batch_index, x1,y1,x2,y2); batch indices round-robin over n_img."""
    b = make_boxes(n, im_h, im_w, seed)
    idx = (np.arange(n) % n_img).astype(np.float64)[:, None]
    return np.hstack([idx, b]).astype(np.float32)


def make_dets(n, im_h=600, im_w=1000, seed=3):
    """f32 [n,5] detections with UNIQUE scores permutation(n)/n (ties would make numpy's
    unstable argsort the arbiter)."""
    b = make_boxes(n, im_h, im_w, seed)
    rng = np.random.default_rng(seed + 101)
    s = (rng.permutation(n).astype(np.float64) + 1.0) / n
    return np.hstack([b, s[:, None]]).astype(np.float32)


class HashNet:
    """A duck-typed AZ 'fc'/'full' net whose outputs are an exactly reproducible integer hash of
    the ROI bits -- no floating-point summation -- so that control-flow parity (levels, dedup,
    decode, clip, unwrap, subdivide, top-N) can be pinned bit-for-bit on any machine.
    Used by oracle/gen_golden.py against the reference's own lib/detect/test.py and by the tests."""

    def __init__(self, seed=11, nsub=11, zoom_rate=0.5, name="hashnet"):
        self.seed, self.nsub, self.zoom_rate, self.name = seed, nsub, zoom_rate, name
        self.inputs = ["conv5_3", "rois"]
        self.outputs = ["zoom_prob", "adj_prob", "adj_bbox"]

        class _B:
            def reshape(self, *s):
                self.shape = s
        self.blobs = {k: _B() for k in ("conv5_3", "rois", "data")}

    @staticmethod
    def _mix(x):
        x = x.astype(np.uint64)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
        return x

    def heads(self, rois):
        """rois f32 [R,5] -> (zoom_prob [R,1], adj_prob [R,nsub], adj_bbox [R,4*nsub]) float32."""
        rois = np.ascontiguousarray(rois, dtype=np.float32)
        bits = rois.view(np.uint32).astype(np.uint64)
        with np.errstate(over="ignore"):
            key = np.full(rois.shape[0], np.uint64(self.seed) * np.uint64(0x9e3779b97f4a7c15), dtype=np.uint64)
            for c in range(5):
                key = self._mix(key ^ (bits[:, c] + np.uint64(0x9e3779b97f4a7c15) * np.uint64(c + 1)))
            ncol = 1 + 5 * self.nsub
            cols = np.arange(ncol, dtype=np.uint64)[None, :]
            h = self._mix(key[:, None] ^ (cols * np.uint64(0xd6e8feb86659fd93)))
        u = ((h >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)  # 24-bit uniform
        zoom = u[:, :1].copy()
        # zoom_prob: u < zoom_rate -> above 0.5
        zoom = np.where(zoom < np.float32(self.zoom_rate), np.float32(0.5) + zoom * np.float32(0.5),
                        zoom * np.float32(0.5)).astype(np.float32)
        adj_prob = u[:, 1:1 + self.nsub].copy()
        adj_bbox = ((u[:, 1 + self.nsub:] - np.float32(0.5)) * np.float32(0.5)).astype(np.float32)
        return zoom, adj_prob, adj_bbox

    def forward(self, blobs=None, **kw):
        z, p, d = self.heads(kw["rois"])
        out = {"zoom_prob": z, "adj_prob": p, "adj_bbox": d}
        for b in (blobs or []):
            out[b] = kw.get(b, np.zeros((1, 1, 1, 1), np.float32))
        return out


class HashDetNet(HashNet):
    """The Fast R-CNN flavour of HashNet: cls_prob [R,C] (24-bit uniforms, not normalised -- the selection only
    compares them) and bbox_pred [R,4C] as an integer hash of the ROI bits.  Pins the detection step's control
    flow (dedup, un-dedup, per-class selection, thresholds, NMS) bit for bit."""

    def __init__(self, seed=13, num_classes=21, name="hashdetnet"):
        HashNet.__init__(self, seed=seed, name=name)
        self.num_classes = num_classes
        self.outputs = ["cls_prob", "bbox_pred"]

    def heads(self, rois):
        rois = np.ascontiguousarray(rois, dtype=np.float32)
        bits = rois.view(np.uint32).astype(np.uint64)
        with np.errstate(over="ignore"):
            key = np.full(rois.shape[0], np.uint64(self.seed) * np.uint64(0x9e3779b97f4a7c15), dtype=np.uint64)
            for c in range(5):
                key = self._mix(key ^ (bits[:, c] + np.uint64(0x9e3779b97f4a7c15) * np.uint64(c + 1)))
            cols = np.arange(5 * self.num_classes, dtype=np.uint64)[None, :]
            h = self._mix(key[:, None] ^ (cols * np.uint64(0xd6e8feb86659fd93)))
        u = ((h >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)
        cls_prob = u[:, :self.num_classes].copy()
        bbox_pred = ((u[:, self.num_classes:] - np.float32(0.5)) * np.float32(0.5)).astype(np.float32)
        return cls_prob, bbox_pred

    def forward(self, blobs=None, **kw):
        p, d = self.heads(kw["rois"])
        out = {"cls_prob": p, "bbox_pred": d}
        for b in (blobs or []):
            out[b] = kw.get(b, np.zeros((1, 1, 1, 1), np.float32))
        return out


class SyntheticImdb:
    """The slice of lib/datasets/imdb.py:16-201 that detect.test touches: image_index,
    image_path_at, name, num_classes, classes, evaluate_detections, competition_mode."""

    def __init__(self, image_paths, num_classes=21, name="synthetic"):
        self._paths = list(image_paths)
        self.image_index = list(range(len(self._paths)))
        self.name = name
        self.classes = ["__background__"] + ["class%d" % i for i in range(1, num_classes)]
        self.num_classes = num_classes
        self.evaluated = None

    def image_path_at(self, i):
        return self._paths[i]

    def evaluate_detections(self, all_boxes, output_dir):
        self.evaluated = (all_boxes, output_dir)

    def competition_mode(self, on):
        pass


class InMemoryImdb(SyntheticImdb):
    """An image database whose images are arrays in host memory.  `image_at(i)` is the optional hook the batched
    drivers use instead of cv2.imread(image_path_at(i)) (aznet_b200/detect/batched.py); everything else is the
    reference's imdb protocol."""

    def __init__(self, images, num_classes=21, name="synthetic_mem"):
        SyntheticImdb.__init__(self, ["<memory:%d>" % i for i in range(len(images))], num_classes, name)
        self._images = list(images)

    def image_at(self, i):
        return self._images[i]
