"""Device-resident Fast R-CNN detection step for a batch of same-sized images (BASELINE config #3).

The B200 restatement of `_frcnn_forward` (lib/detect/test.py:259-318) and of the per-class selection of
`test_net` (:549-553, :608-651) + `apply_nms` (:467-484), batched over images AND classes:

    azn_detect_rois     proposals -> ROI blob, feature-space dedup per image / chunk
    azn_roi_pool_fwd    ROI max-pool of the unique ROIs over the cached conv5_3 maps (staged kernel)
    azn_fc_forward x3   fc6(+ReLU) -> fc7(+ReLU) -> [cls_score | bbox_pred] (+softmax on the class columns)
    azn_detect_select   un-dedup, score > thresh[j], top-100 per (image, class), decode + clip of the winners
then, once per image set (after an all-gather of the score tensor on several GPUs):
    azn_detect_thresholds / azn_detect_filter / azn_nms_segments

Nothing leaves HBM between the proposals and the kept detections; the reference round-trips to the host per
image and runs 80 Python heap loops and 80 Cython NMS calls per image.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import ops
from .engine import im_scale_for



# ROI-pool kernel of the detection step: "staged" (shared-memory-staged kernel over per-image buckets: ~300 proposals per
# image), "direct" (the L2-fed kernel, one CTA per bin row) or "per_roi" (the L2-fed kernel, one CTA per ROI): same bits,
# an A/B switch for benchmarks (tools/detbench.py --pool).
POOL_KERNEL = "staged"


def pool_kwargs():
    return {"staged": dict(staged=True), "direct": dict(staged=False), "per_roi": dict(per_roi=True)}[POOL_KERNEL]

class DetectEngine:
    """Buffers + launch sequence of the batched detection step.  One instance per (n_img, image shape, cap_boxes)."""

    def __init__(self, head, n_img, im_h, im_w, cap_boxes=300, *, scales=(600,), max_size=1000, batch_size=10000,
                 dedup=1. / 16., eps=1e-14, max_per_image=100, spatial_scale=0.0625, device=None):
        L.require_device()
        self.head = head
        self.dev = device or head.w6.device
        self.n_img, self.im_h, self.im_w, self.cap = int(n_img), int(im_h), int(im_w), int(cap_boxes)
        self.C = int(head.num_classes)
        self.mpi = int(max_per_image)
        self.scale = im_scale_for(im_h, im_w, scales, max_size)
        self.spatial_scale = float(spatial_scale)
        dev, i32 = self.dev, torch.int32
        n, cap, Cc, mpi = self.n_img, self.cap, self.C, self.mpi
        z = lambda *s, dt=i32: torch.zeros(s, dtype=dt, device=dev)
        self.im_h_d = torch.full((n,), self.im_h, dtype=i32, device=dev)
        self.im_w_d = torch.full((n,), self.im_w, dtype=i32, device=dev)
        self.im_scale_d = torch.full((n,), self.scale, dtype=torch.float64, device=dev)
        self.inv, self.rep, self.n_uniq, self.img_off = z(n, cap), z(n, cap), z(n), z(n + 1)
        self.rois = z(n * cap, 5, dt=torch.float32)
        self.m_total = z(1)
        self.hashes, self.flags = z(n, cap, dt=torch.int64), z(n, cap)
        m_cap = n * cap
        k6 = head.w6.shape[1]
        self.pool5 = torch.empty((m_cap, k6), dtype=torch.bfloat16, device=dev)
        self.h6 = torch.empty((m_cap, head.w6.shape[0]), dtype=torch.bfloat16, device=dev)
        self.h7 = torch.empty((m_cap, head.w7.shape[0]), dtype=torch.bfloat16, device=dev)
        # skip-layer head: three pooled-position matrices, their normalised concat, the live row count
        self.skip = hasattr(head, "conv_names")
        if self.skip:
            P2 = head.pooled * head.pooled
            self.cat = torch.zeros((m_cap * P2, head.k_cat), dtype=torch.bfloat16, device=dev)
            self.m_rows = z(1)
        self.n_out = 5 * Cc
        self.ld = (self.n_out + 3) // 4 * 4
        self.out = torch.zeros((m_cap, self.ld), dtype=torch.float32, device=dev)
        # results of one batch (callers that stream many batches pass slices of their own set-wide buffers)
        self.dets = z(n, Cc, mpi, 5, dt=torch.float32)
        self.top_scores = torch.full((n, Cc, mpi), float("-inf"), dtype=torch.float32, device=dev)
        self.det_count = z(n, Cc)
        self.launches = 0
        st = self._st = L.DetectState()
        p = lambda t: t.data_ptr()
        st.n_img, st.cap_boxes, st.num_classes, st.max_per_image = n, cap, Cc, mpi
        st.chunk, st.ld_head = int(batch_size), self.ld
        st.im_h, st.im_w, st.im_scale = p(self.im_h_d), p(self.im_w_d), p(self.im_scale_d)
        st.eps, st.dedup = float(eps), float(dedup)
        st.inv, st.rep, st.n_uniq, st.img_off = p(self.inv), p(self.rep), p(self.n_uniq), p(self.img_off)
        st.rois, st.m_total, st.hashes, st.flags = p(self.rois), p(self.m_total), p(self.hashes), p(self.flags)
        st.head_out = p(self.out)
        st.thresh = None

    def prepare(self, boxes: torch.Tensor, counts: torch.Tensor):
        """Proposals -> deduplicated ROI blob (self.rois[:m_total], self.inv / rep / img_off)."""
        st = self._st
        assert boxes.dtype == torch.float64 and boxes.is_contiguous() and tuple(boxes.shape) == (self.n_img, self.cap, 4)
        assert counts.dtype == torch.int32 and counts.numel() == self.n_img and counts.is_contiguous()
        self._boxes, self._counts = boxes, counts           # azn_detect_select reads them again: keep the storage alive
        st.boxes, st.n_boxes = boxes.data_ptr(), counts.data_ptr()
        L.check(L.lib().azn_detect_rois(C.byref(st), ops._stream()), "azn_detect_rois")
        self.launches += 2

    def _skip_pool5(self, maps: dict):
        """roi_pool{3,4,5} with GRN + concat + x1000 fused (azn_roi_pool_grn_fwd) -> conv_pool5 (+ReLU) into self.pool5
        (VGG16_skip test_fc.prototxt:28-142).  The pooled intermediates never reach HBM."""
        hd = self.head
        mc, P = self.n_img * self.cap, hd.pooled
        off = 0
        for name, sc, c in zip(hd.conv_names, hd.scales, hd.src_channels):
            m = maps[name]
            assert m.dtype == torch.bfloat16 and m.shape[0] == self.n_img and m.is_contiguous()
            ops.roi_pool_grn(m, self.rois, self.cat, off, P, sc, hd.grn_scale, n_rois=self.m_total)
            off += c
        torch.mul(self.m_total, P * P, out=self.m_rows)
        ops.fc_forward(self.cat, hd.wc, hd.bc, L.ACT_RELU, m_live=self.m_rows, out=self.pool5.view(mc * P * P, hd.C))
        self.launches += 3 + 1 + 2
        return self.pool5

    def run_head(self, conv_nhwc):
        """ROI pool (staged kernel: ~300 ROIs per image) + fc6 / fc7 / [cls_score | bbox_pred] with softmax.
        conv_nhwc: the bf16 NHWC conv5_3 maps, or for the skip-layer head a dict name -> maps."""
        hd = self.head
        mc = self.n_img * self.cap
        if self.skip:
            pool = self._skip_pool5(conv_nhwc)
        else:
            assert conv_nhwc.dtype == torch.bfloat16 and conv_nhwc.shape[0] == self.n_img and conv_nhwc.is_contiguous()
            pool = ops.roi_pool(conv_nhwc, self.rois, hd.pooled, self.spatial_scale, layout="NHWC", n_rois=self.m_total,
                                out=self.pool5.view(mc, hd.pooled, hd.pooled, hd.C), **pool_kwargs())
        ops.fc_forward(pool.view(mc, -1), hd.w6, hd.b6, L.ACT_RELU, m_live=self.m_total, out=self.h6)
        ops.fc_forward(self.h6, hd.w7, hd.b7, L.ACT_RELU, m_live=self.m_total, out=self.h7)
        ops.fc_forward(self.h7, hd.wo, hd.bo, L.ACT_SOFTMAX_BBOX, self.C, m_live=self.m_total, out=self.out[:, :self.n_out])
        self.launches += 3 + 3 + 1

    def select(self, *, thresh=None, dets=None, top_scores=None, det_count=None, head_out=None):
        """Per (image, class): score > thresh[j], top-100, decode + clip -> dets / top_scores / det_count (own
        buffers unless given).  `head_out` f32 [n*cap, ld] replaces the net's output (tests)."""
        st = self._st
        dets = self.dets if dets is None else dets
        top_scores = self.top_scores if top_scores is None else top_scores
        det_count = self.det_count if det_count is None else det_count
        for t, shape in ((dets, (self.n_img, self.C, self.mpi, 5)), (top_scores, (self.n_img, self.C, self.mpi)),
                         (det_count, (self.n_img, self.C))):
            assert t.is_contiguous() and tuple(t.shape) == shape
        st.dets, st.top_scores, st.det_count = dets.data_ptr(), top_scores.data_ptr(), det_count.data_ptr()
        st.thresh = thresh.data_ptr() if thresh is not None else None
        st.head_out = (self.out if head_out is None else head_out).data_ptr()
        L.check(L.lib().azn_detect_select(C.byref(st), ops._stream()), "azn_detect_select")
        self.launches += 1
        return dets, top_scores, det_count

    def detect(self, conv_nhwc: torch.Tensor, boxes: torch.Tensor, counts: torch.Tensor, **out):
        """conv_nhwc bf16 [n,H,W,C]; boxes f64 [n, cap_boxes, 4] proposals (image coordinates), counts int32 [n].
        Asynchronous.  Returns (dets f32 [n,C,100,5], top_scores f32 [n,C,100], det_count int32 [n,C])."""
        self.prepare(boxes, counts)
        self.run_head(conv_nhwc)
        return self.select(**out)


class DetectionSet:
    """Set-wide device buffers of a detection run: every batch's azn_detect_select writes into its slice, the
    thresholds / filter / NMS run once at the end (`finish`).  On several GPUs every rank holds its shard of the
    images; `finish` all-gathers the [images, C, 100] score tensor so that every rank computes the same
    thresholds (the one exchange step of the detection path), then filters and suppresses locally."""

    def __init__(self, num_images, num_classes, max_per_image=100, device=None, total_images=None):
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.N, self.C, self.mpi = int(num_images), int(num_classes), int(max_per_image)
        self.total_images = int(total_images) if total_images is not None else self.N
        self.dets = torch.zeros((self.N, self.C, self.mpi, 5), dtype=torch.float32, device=dev)
        self.top_scores = torch.full((self.N, self.C, self.mpi), float("-inf"), dtype=torch.float32, device=dev)
        self.det_count = torch.zeros((self.N, self.C), dtype=torch.int32, device=dev)
        self.thresh = self.keep = self.keep_count = None

    def slot(self, lo, hi):
        return dict(dets=self.dets[lo:hi], top_scores=self.top_scores[lo:hi], det_count=self.det_count[lo:hi])

    @property
    def max_per_set(self):
        return 800 // (self.C - 1) * self.total_images          # Python-2 integer division (test.py:551)

    def finish(self, nms_thresh):
        from .dist import gather_detection_scores
        top, cnt = gather_detection_scores(self.top_scores, self.det_count)
        self.thresh = ops.detect_thresholds(top, cnt, self.max_per_set)
        _, self.keep, self.keep_count = finish_detections(self.dets, self.top_scores, self.det_count, self.max_per_set,
                                                          nms_thresh, thresh=self.thresh)
        return self.thresh, self.keep, self.keep_count

    def to_host(self, nms=False):
        return detections_to_host(self.dets, self.det_count, self.keep if nms else None, self.keep_count if nms else None)


def finish_detections(dets: torch.Tensor, top_scores: torch.Tensor, det_count: torch.Tensor, max_per_set: int,
                      nms_thresh: float, thresh: torch.Tensor | None = None):
    """The end of test_net for a whole image set (lib/detect/test.py:624-651 + apply_nms :467-484), on the device:
    per-class thresholds from all pushed scores, the final strict filter (det_count shrinks in place), NMS of
    every (image, class) problem in one launch.  dets [N,C,mpi,5], top_scores [N,C,mpi], det_count [N,C].
    Returns (thresh f32 [C], keep int64 [N,C,mpi] positions inside the (image, class) rows, keep_count int32 [N,C])."""
    N, Cc, mpi = top_scores.shape
    if thresh is None:
        thresh = ops.detect_thresholds(top_scores, det_count, max_per_set)
    ops.detect_filter(top_scores, det_count, thresh)
    seg_off = torch.arange(N * Cc, dtype=torch.int32, device=dets.device) * mpi
    keep, keep_count = ops.nms_segments(dets.view(-1, 5), seg_off, det_count.view(-1), mpi, float(nms_thresh))
    return thresh, keep.view(N, Cc, mpi), keep_count.view(N, Cc)


def detections_to_host(dets, det_count, keep=None, keep_count=None):
    """Device results -> the reference's nesting all_boxes[cls][img] of float32 [n,5] arrays ([] for class 0 and
    for empty NMS results, like apply_nms).  With keep lists: the post-NMS detections in keep order."""
    d, c = dets.cpu().numpy(), det_count.cpu().numpy()
    N, Cc = c.shape
    if keep is not None:
        k, kc = keep.cpu().numpy(), keep_count.cpu().numpy()
    out = [[[] for _ in range(N)] for _ in range(Cc)]
    for j in range(1, Cc):
        for i in range(N):
            if keep is None:
                out[j][i] = d[i, j, :c[i, j]].copy()
            elif kc[i, j] > 0:
                out[j][i] = d[i, j, k[i, j, :kc[i, j]]].copy()
    return out
