"""Torch-tensor front end of the C ABI (include/aznet_b200.h).

PyTorch is used for device memory and streams only; every function here ends in exactly one
call into libaznet_b200.so on torch's current stream.  Nothing falls back to PyTorch math.
"""
from __future__ import annotations

import torch

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("aznet_b200 ops need CUDA tensors (no CPU fallback)")


def roi_pool(feat: torch.Tensor, rois: torch.Tensor, pooled: int = 7, spatial_scale: float = 0.0625,
             layout: str = "NCHW", n_rois: torch.Tensor | None = None, want_argmax: bool = False,
             out: torch.Tensor | None = None, staged: bool | None = None, per_roi: bool = False):
    """ROI max pooling (roi_pooling_layer.cpp:46-125).  feat [n,C,H,W] (NCHW) or [n,H,W,C] (NHWC),
    f32 or bf16; rois f32 [R,5].  Returns pooled [R,C,P,P] (NCHW) or [R,P,P,C] (NHWC)."""
    _need_cuda(feat, rois, n_rois)
    assert feat.is_contiguous() and rois.is_contiguous() and rois.dtype == torch.float32
    if layout == "NCHW":
        n, Cc, H, W = feat.shape
    else:
        n, H, W, Cc = feat.shape
    R = rois.shape[0]
    dt = {torch.float32: L.DTYPE_F32, torch.bfloat16: L.DTYPE_BF16}[feat.dtype]
    shape = (R, Cc, pooled, pooled) if layout == "NCHW" else (R, pooled, pooled, Cc)
    if out is None:
        out = torch.empty(shape, dtype=feat.dtype, device=feat.device)
    amax = torch.empty(shape, dtype=torch.int32, device=feat.device) if want_argmax else None
    lay = L.LAYOUT_NCHW if layout == "NCHW" else L.LAYOUT_NHWC
    nbytes = L.lib().azn_roi_pool_workspace_bytes(n, Cc, H, W, lay, dt, R)
    ws = _scratch(feat.device, nbytes) if nbytes else None
    if staged is not None or per_roi:          # many ROIs per image with a device-side count: ask for the staged kernel;
        # per_roi: many SMALL ROIs (deep search levels): the direct kernel with one CTA per ROI (NHWC)
        choice = 3 if (per_roi and layout == "NHWC") else (2 if staged else 1)
        L.check(L.lib().azn_roi_pool_fwd_ex(_ptr(feat), n, Cc, H, W, lay, dt, _ptr(rois), _ptr(n_rois), R, pooled, pooled,
                                            spatial_scale, _ptr(out), _ptr(amax), _ptr(ws), nbytes, choice, _stream()),
                "azn_roi_pool_fwd_ex")
    else:
        L.check(L.lib().azn_roi_pool_fwd(_ptr(feat), n, Cc, H, W, lay, dt, _ptr(rois), _ptr(n_rois), R, pooled, pooled,
                                         spatial_scale, _ptr(out), _ptr(amax), _ptr(ws), nbytes, _stream()), "azn_roi_pool_fwd")
    return (out, amax) if want_argmax else out


def nchw_to_nhwc_bf16(src: torch.Tensor, out: torch.Tensor | None = None):
    _need_cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous()
    n, Cc, H, W = src.shape
    if out is None:
        out = torch.empty((n, H, W, Cc), dtype=torch.bfloat16, device=src.device)
    L.check(L.lib().azn_nchw_f32_to_nhwc_bf16(_ptr(src), n, Cc, H, W, _ptr(out), _stream()), "azn_nchw_f32_to_nhwc_bf16")
    return out


def nchw_bf16_to_nhwc_bf16(src: torch.Tensor, out: torch.Tensor | None = None):
    """bf16 NCHW (a batch narrowed on the host, `host_f32_to_bf16`) -> the engine's bf16 NHWC."""
    _need_cuda(src)
    assert src.dtype == torch.bfloat16 and src.is_contiguous()
    n, Cc, H, W = src.shape
    if out is None:
        out = torch.empty((n, H, W, Cc), dtype=torch.bfloat16, device=src.device)
    L.check(L.lib().azn_nchw_bf16_to_nhwc_bf16(_ptr(src), n, Cc, H, W, _ptr(out), _stream()), "azn_nchw_bf16_to_nhwc_bf16")
    return out


def host_f32_to_bf16(src: torch.Tensor, out: torch.Tensor, threads: int = 0):
    """HOST tensors: out (bf16, same element count, may be pinned) = round-to-nearest-even of src (f32), on `threads`
    worker threads of the library (0: all hardware threads).  The call releases the GIL (ctypes)."""
    assert src.device.type == "cpu" and out.device.type == "cpu", "host_f32_to_bf16 converts host arrays"
    assert src.dtype == torch.float32 and out.dtype == torch.bfloat16 and src.is_contiguous() and out.is_contiguous()
    assert src.numel() == out.numel()
    L.check(L.lib().azn_host_f32_to_bf16(src.data_ptr(), out.data_ptr(), src.numel(), int(threads)), "azn_host_f32_to_bf16")
    return out


_WS = {}
_SCRATCH = {}


def _scratch(device, nbytes):
    """Grow-only scratch buffer per (device, stream): reuse is ordered by the stream it belongs to."""
    key = (device, torch.cuda.current_stream().cuda_stream)
    t = _SCRATCH.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _SCRATCH[key] = t
    return t


def fc_workspace(device, N):
    """The fixed stream-K workspace (one 128 x BLOCK_N fp32 slot per SM + flags), zeroed once, cached per
    (device, stream)."""
    nbytes = L.lib().azn_fc_workspace_bytes(1, int(N), 64)
    key = (device, nbytes, torch.cuda.current_stream().cuda_stream)
    if key not in _WS:
        _WS[key] = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    return _WS[key]


def fc_forward(A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, act: int = L.ACT_NONE, act_aux: int = 0,
               out_dtype=torch.bfloat16, m_live: torch.Tensor | None = None, out: torch.Tensor | None = None):
    """out = act(A . W^T + bias) on the tensor cores (inner_product_layer.cpp:80-93).  A bf16 [M,K],
    W bf16 [N,K], bias f32 [N]."""
    _need_cuda(A, W, bias, m_live)
    assert A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16 and bias.dtype == torch.float32
    assert A.is_contiguous() and W.is_contiguous() and A.shape[1] == W.shape[1]
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=A.device)
    ws = fc_workspace(A.device, N)
    od = {torch.float32: L.DTYPE_F32, torch.bfloat16: L.DTYPE_BF16}[out.dtype]
    L.check(L.lib().azn_fc_forward(_ptr(A), _ptr(W), _ptr(bias), _ptr(out), od, out.stride(0), M, _ptr(m_live), N, K,
                                   act, act_aux, _ptr(ws), ws.numel(), _stream()), "azn_fc_forward")
    return out


def az_heads(h7: torch.Tensor, wh: torch.Tensor, bias: torch.Tensor, nsub: int, m_live: torch.Tensor | None = None,
             out: torch.Tensor | None = None):
    """[adj_prob | adj_bbox | zoom_prob] = the AZ head's three output layers + sigmoids (test_fc.prototxt:146-232) in
    one small mma.sync kernel (azn_az_heads_forward).  h7 bf16 [M, K], wh bf16 [N = 5*nsub+1, K], bias f32 [N];
    out f32 [M, ld >= N]."""
    _need_cuda(h7, wh, bias, m_live)
    assert h7.dtype == torch.bfloat16 and wh.dtype == torch.bfloat16 and bias.dtype == torch.float32
    assert h7.is_contiguous() and wh.is_contiguous() and h7.shape[1] == wh.shape[1]
    M, K = h7.shape
    N = wh.shape[0]
    if out is None:
        out = torch.empty((M, (N + 7) // 8 * 8), dtype=torch.float32, device=h7.device)
    assert out.dtype == torch.float32 and out.shape[0] >= M and out.stride(1) == 1
    L.check(L.lib().azn_az_heads_forward(_ptr(h7), _ptr(wh), _ptr(bias), _ptr(out), out.stride(0), M, _ptr(m_live), N, K, int(nsub),
                                         _stream()), "azn_az_heads_forward")
    return out


def nms(dets: torch.Tensor, thresh: float):
    """Greedy NMS (lib/utils/nms.pyx:17-68) on a CUDA f32 [n,5] tensor.  Returns (keep int64 [n],
    count int32 [1]) on the device; keep[:count] are the kept indices in descending score order."""
    _need_cuda(dets)
    assert dets.dtype == torch.float32 and dets.dim() == 2 and dets.shape[1] == 5 and dets.is_contiguous()
    n = dets.shape[0]
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=dets.device)
    cnt = torch.empty(1, dtype=torch.int32, device=dets.device)          # always written by azn_nms (a memset when n == 0)
    nbytes = L.lib().azn_nms_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dets.device)
    L.check(L.lib().azn_nms(_ptr(dets), n, float(thresh), _ptr(keep), _ptr(cnt), _ptr(ws), nbytes, _stream()), "azn_nms")
    return keep, cnt


def nms_batched(dets: torch.Tensor, seg_off: torch.Tensor, thresh: float):
    """Independent NMS problems in one launch.  dets f32 [T,5], seg_off int32 [S+1] (device).
    Returns (keep int64 [T] segment-local indices at seg_off[s], counts int32 [S])."""
    _need_cuda(dets, seg_off)
    assert dets.dtype == torch.float32 and seg_off.dtype == torch.int32 and dets.is_contiguous()
    S = seg_off.numel() - 1
    keep = torch.empty(max(dets.shape[0], 1), dtype=torch.int64, device=dets.device)
    cnt = torch.zeros(max(S, 1), dtype=torch.int32, device=dets.device)
    L.check(L.lib().azn_nms_batched(_ptr(dets), _ptr(seg_off), S, float(thresh), _ptr(keep), _ptr(cnt), _stream()),
            "azn_nms_batched")
    return keep, cnt[:S]


def nms_segments(dets: torch.Tensor, seg_off: torch.Tensor, seg_len: torch.Tensor, max_len: int, thresh: float):
    """NMS over padded storage: segment s = rows [seg_off[s], seg_off[s] + seg_len[s]) of dets f32 [T,5].
    Returns (keep int64 [T] segment-local indices at seg_off[s], counts int32 [S])."""
    _need_cuda(dets, seg_off, seg_len)
    assert dets.dtype == torch.float32 and seg_off.dtype == torch.int32 and seg_len.dtype == torch.int32
    assert dets.is_contiguous() and seg_off.is_contiguous() and seg_len.is_contiguous()
    S = seg_off.numel()
    keep = torch.empty(max(dets.shape[0], 1), dtype=torch.int64, device=dets.device)
    cnt = torch.zeros(max(S, 1), dtype=torch.int32, device=dets.device)
    L.check(L.lib().azn_nms_segments(_ptr(dets), _ptr(seg_off), _ptr(seg_len), S, int(max_len), float(thresh), _ptr(keep),
                                     _ptr(cnt), _stream()), "azn_nms_segments")
    return keep, cnt[:S]


def detect_thresholds(top_scores: torch.Tensor, det_count: torch.Tensor, max_per_set: int, out: torch.Tensor | None = None):
    """test_net's per-class thresholds after the whole image set (lib/detect/test.py:624-631).
    top_scores f32 [n, C, mpi] score-descending rows, det_count int32 [n, C].  Returns f32 [C]."""
    _need_cuda(top_scores, det_count)
    assert top_scores.dtype == torch.float32 and det_count.dtype == torch.int32
    assert top_scores.is_contiguous() and det_count.is_contiguous()
    n, Cc, mpi = top_scores.shape
    if out is None:
        out = torch.empty(Cc, dtype=torch.float32, device=top_scores.device)
    L.check(L.lib().azn_detect_thresholds(_ptr(top_scores), _ptr(det_count), n, Cc, mpi, int(max_per_set), _ptr(out),
                                          _stream()), "azn_detect_thresholds")
    return out


def roi_pool_grn(feat: torch.Tensor, rois: torch.Tensor, out: torch.Tensor, ch_off: int, pooled: int = 7,
                 spatial_scale: float = 0.0625, grn_scale: float = 1000.0, n_rois: torch.Tensor | None = None):
    """ROI max pooling with the skip head's GRN + concat + Power epilogue fused (azn_roi_pool_grn_fwd).
    feat bf16 [n,H,W,C]; out bf16 [R*P*P, ld]: columns [ch_off, ch_off+C) are written."""
    _need_cuda(feat, rois, out, n_rois)
    assert feat.dtype == torch.bfloat16 and feat.is_contiguous() and rois.is_contiguous() and rois.dtype == torch.float32
    n, H, W, Cc = feat.shape
    R = rois.shape[0]
    assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.dim() == 2 and out.shape[0] >= R * pooled * pooled
    L.check(L.lib().azn_roi_pool_grn_fwd(_ptr(feat), n, Cc, H, W, _ptr(rois), _ptr(n_rois), R, pooled, pooled, float(spatial_scale),
                                         float(grn_scale), _ptr(out), int(out.shape[1]), int(ch_off), _stream()), "azn_roi_pool_grn_fwd")
    return out


def grn_concat(pooled: list, scale: float = 1000.0, n_units: torch.Tensor | None = None, rows_per_unit: int = 49,
               out: torch.Tensor | None = None):
    """GRN of every source + channel concat + Power scale (VGG16_skip test_fc.prototxt:39-110, grn_layer.cpp:27-56).
    pooled: list of bf16 [rows, C_l] (pooled positions x channels).  Returns bf16 [rows, sum C_l]."""
    import ctypes as C
    _need_cuda(*pooled, n_units)
    rows = pooled[0].shape[0]
    for t in pooled:
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.dim() == 2 and t.shape[0] == rows
    ctot = sum(int(t.shape[1]) for t in pooled)
    if out is None:
        out = torch.empty((rows, ctot), dtype=torch.bfloat16, device=pooled[0].device)
    assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.shape[0] == rows and out.shape[1] >= ctot
    ptrs = (C.c_void_p * len(pooled))(*[t.data_ptr() for t in pooled])
    chans = (C.c_int32 * len(pooled))(*[int(t.shape[1]) for t in pooled])
    L.check(L.lib().azn_grn_concat_forward(ptrs, chans, len(pooled), _ptr(n_units), rows, int(rows_per_unit), float(scale),
                                           _ptr(out), int(out.shape[1]), _stream()), "azn_grn_concat_forward")
    return out


def tune_threshold(zoom: torch.Tensor, counts: torch.Tensor, max_per_set: int):
    """tune_thresh's zoom threshold (lib/detect/tune.py:318-366): the max_per_set-th highest anchor zoom score of the
    image set, -inf when fewer anchors were seen.  zoom f32 [n, cap], counts int32 [n].  Returns f32 [1]."""
    _need_cuda(zoom, counts)
    assert zoom.dtype == torch.float32 and counts.dtype == torch.int32 and zoom.is_contiguous() and counts.is_contiguous()
    n, cap = zoom.shape
    out = torch.empty(1, dtype=torch.float32, device=zoom.device)
    L.check(L.lib().azn_tune_threshold(_ptr(zoom), _ptr(counts), n, cap, int(max_per_set), _ptr(out), _stream()),
            "azn_tune_threshold")
    return out


def detect_filter(top_scores: torch.Tensor, det_count: torch.Tensor, thresh: torch.Tensor):
    """Final `score > thresh[class]` filter of test_net (:646-651): shrinks det_count in place."""
    _need_cuda(top_scores, det_count, thresh)
    assert thresh.dtype == torch.float32 and det_count.dtype == torch.int32 and det_count.is_contiguous()
    n, Cc, mpi = top_scores.shape
    L.check(L.lib().azn_detect_filter(_ptr(top_scores), _ptr(det_count), _ptr(thresh), n, Cc, mpi, _stream()),
            "azn_detect_filter")
    return det_count


def divide_region(regions: torch.Tensor, min_side: float, sift_only: bool = False):
    """divide_region + _sift_dup (lib/utils/div.pyx:15-88) for one region set, f64 [n,4] on the device.
    Returns (out f64 [cap,4], count int32 [1])."""
    _need_cuda(regions)
    assert regions.dtype == torch.float64 and regions.is_contiguous()
    n = regions.shape[0]
    cap = max(n * 16 + 64, 1)
    out = torch.empty((cap, 4), dtype=torch.float64, device=regions.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=regions.device)
    nbytes = L.lib().azn_divide_region_scratch_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=regions.device)
    L.check(L.lib().azn_divide_region(_ptr(regions), n, float(min_side), _ptr(out), _ptr(cnt), cap, int(sift_only),
                                      _ptr(ws), nbytes, _stream()), "azn_divide_region")
    return out, cnt


def decode_boxes(boxes: torch.Tensor, deltas: torch.Tensor, im_h: int, im_w: int, eps: float = 1e-14):
    """_bbox_pred + _clip_boxes (lib/detect/test.py:106-151).  boxes f64 [n,4], deltas f32 [n,4k]."""
    _need_cuda(boxes, deltas)
    assert boxes.dtype == torch.float64 and deltas.dtype == torch.float32
    boxes, deltas = boxes.contiguous(), deltas.contiguous()
    n, ncol = boxes.shape[0], deltas.shape[1] // 4
    out = torch.empty((n, 4 * ncol), dtype=torch.float64, device=boxes.device)
    L.check(L.lib().azn_decode_boxes(_ptr(boxes), _ptr(deltas), n, ncol, float(eps), int(im_h), int(im_w), _ptr(out),
                                     _stream()), "azn_decode_boxes")
    return out


# ---- the conv5_3 backbone: zero-bordered channels-last maps [n, H+2, W+2, C] bf16 -----------------
def blob_size(h0: int, w0: int, im_scale: float):
    """Destination size of cv2.resize(im, None, None, fx=s, fy=s): cvRound (half to even) of size * s."""
    import numpy as np
    return int(np.rint(h0 * im_scale)), int(np.rint(w0 * im_scale))


def image_blob(images: torch.Tensor, im_scale: float, pixel_means, cpad: int = 64, out: torch.Tensor | None = None,
               want_f32: bool = False):
    """uint8 [n, H0, W0, 3] BGR images -> mean-subtracted, bilinearly resized network input
    (_get_image_blob, lib/detect/test.py:27-59) as a zero-bordered NHWC bf16 map [n, Hs+2, Ws+2, cpad]
    (+ Caffe's f32 NCHW 'data' blob [n, 3, Hs, Ws] when want_f32)."""
    _need_cuda(images)
    assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[3] == 3 and images.is_contiguous()
    n, h0, w0, _ = images.shape
    hs, ws = blob_size(h0, w0, im_scale)
    if out is None:
        out = torch.empty((n, hs + 2, ws + 2, cpad), dtype=torch.bfloat16, device=images.device)
    assert out.shape == (n, hs + 2, ws + 2, cpad) and out.dtype == torch.bfloat16 and out.is_contiguous()
    blob = torch.empty((n, 3, hs, ws), dtype=torch.float32, device=images.device) if want_f32 else None
    import ctypes as C
    means = (C.c_double * 3)(*[float(m) for m in pixel_means])
    L.check(L.lib().azn_image_blob(_ptr(images), n, h0, w0, float(im_scale), means, _ptr(out), cpad, hs, ws, _ptr(blob),
                                   _stream()), "azn_image_blob")
    return (out, blob) if want_f32 else out


def pack_conv_weight(w: torch.Tensor, cin_pad: int | None = None):
    """Caffe conv weight [Cout, Cin, 3, 3] -> bf16 [Cout, 9 * Cin_pad] with column (ky*3+kx)*Cin_pad + c
    (the K order of azn_conv3x3_forward); input channels zero-padded to a multiple of 64."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    cp = cin_pad or (ci + 63) // 64 * 64
    t = torch.zeros((co, 3, 3, cp), dtype=torch.float32, device=w.device)
    t[..., :ci] = w.permute(0, 2, 3, 1)
    return t.reshape(co, 9 * cp).to(torch.bfloat16).contiguous()


def conv3x3(x: torch.Tensor, wt: torch.Tensor, bias: torch.Tensor, relu: bool = True, out: torch.Tensor | None = None,
            unpadded: bool = False):
    """3x3 / pad 1 / stride 1 convolution (+ bias, ReLU) as an implicit GEMM on the tensor cores.
    x bf16 [n, H+2, W+2, Cin] zero-bordered; wt = pack_conv_weight(W); -> [n, H+2, W+2, Cout] zero-bordered, or
    [n, H, W, Cout] with unpadded=True."""
    _need_cuda(x, wt, bias)
    assert x.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16 and bias.dtype == torch.float32
    assert x.is_contiguous() and wt.is_contiguous() and x.dim() == 4
    n, hp, wp, cin = x.shape
    cout = wt.shape[0]
    assert wt.shape[1] == 9 * cin
    shape = (n, hp - 2, wp - 2, cout) if unpadded else (n, hp, wp, cout)
    if out is None:
        out = torch.empty(shape, dtype=torch.bfloat16, device=x.device)
    assert tuple(out.shape) == shape and out.is_contiguous() and out.dtype == torch.bfloat16
    ws = fc_workspace(x.device, cout)
    L.check(L.lib().azn_conv3x3_forward(_ptr(x), _ptr(wt), _ptr(bias), _ptr(out), n, hp - 2, wp - 2, cin, cout, int(relu),
                                        int(unpadded), _ptr(ws), ws.numel(), _stream()), "azn_conv3x3_forward")
    return out


def patches3x3(x: torch.Tensor, cin: int, kp: int = 64, out: torch.Tensor | None = None):
    """Zero-bordered [n, H+2, W+2, Cs] bf16 (Cs >= cin) -> [n, H+2, W+2, kp]: every pixel's 3x3 x cin neighbourhood as one
    K row, entry (ky*3+kx)*cin + c (azn_patches3x3) -- the operand of conv_patches."""
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    n, hp, wp, cs = x.shape
    if out is None:
        out = torch.empty((n, hp, wp, kp), dtype=torch.bfloat16, device=x.device)
    assert tuple(out.shape) == (n, hp, wp, kp) and out.is_contiguous() and out.dtype == torch.bfloat16
    L.check(L.lib().azn_patches3x3(_ptr(x), n, hp - 2, wp - 2, cs, int(cin), _ptr(out), kp, _stream()), "azn_patches3x3")
    return out


def conv_direct(x: torch.Tensor, cin: int, wt: torch.Tensor, bias: torch.Tensor, relu: bool = True, out: torch.Tensor | None = None):
    """The first 3x3 convolution in one kernel (azn_conv3x3_direct_forward): x zero-bordered [n, H+2, W+2, Cs] bf16 with
    9 * cin <= 32, wt = pack_patch_weight(W) [64, Kp] -> zero-bordered [n, H+2, W+2, 64]."""
    _need_cuda(x, wt, bias)
    assert x.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16 and bias.dtype == torch.float32
    assert x.is_contiguous() and wt.is_contiguous() and x.dim() == 4
    n, hp, wp, cs = x.shape
    cout, kp = wt.shape
    if out is None:
        out = torch.empty((n, hp, wp, cout), dtype=torch.bfloat16, device=x.device)
    assert tuple(out.shape) == (n, hp, wp, cout) and out.is_contiguous() and out.dtype == torch.bfloat16
    L.check(L.lib().azn_conv3x3_direct_forward(_ptr(x), n, hp - 2, wp - 2, cs, int(cin), _ptr(wt), kp, _ptr(bias), _ptr(out), cout,
                                               int(relu), _stream()), "azn_conv3x3_direct_forward")
    return out


def pack_patch_weight(w: torch.Tensor, kp: int = 64):
    """Caffe conv weight [Cout, Cin, 3, 3] with 9*Cin <= kp -> bf16 [Cout, kp], column (ky*3+kx)*Cin + c, zero-padded."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3 and 9 * ci <= kp
    t = torch.zeros((co, kp), dtype=torch.float32, device=w.device)
    t[:, :9 * ci] = w.permute(0, 2, 3, 1).reshape(co, 9 * ci)
    return t.to(torch.bfloat16).contiguous()


def conv_patches(xp: torch.Tensor, wt: torch.Tensor, bias: torch.Tensor, relu: bool = True, out: torch.Tensor | None = None,
                 unpadded: bool = False):
    """The 3x3 convolution over gathered patches (azn_conv_patches_forward): xp bf16 [n, H+2, W+2, Kp] from patches3x3,
    wt = pack_patch_weight(W) -> zero-bordered [n, H+2, W+2, Cout] (or [n, H, W, Cout])."""
    _need_cuda(xp, wt, bias)
    assert xp.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16 and bias.dtype == torch.float32
    assert xp.is_contiguous() and wt.is_contiguous() and xp.dim() == 4 and wt.shape[1] == xp.shape[3]
    n, hp, wp, kp = xp.shape
    cout = wt.shape[0]
    shape = (n, hp - 2, wp - 2, cout) if unpadded else (n, hp, wp, cout)
    if out is None:
        out = torch.empty(shape, dtype=torch.bfloat16, device=xp.device)
    assert tuple(out.shape) == shape and out.is_contiguous() and out.dtype == torch.bfloat16
    ws = fc_workspace(xp.device, cout)
    L.check(L.lib().azn_conv_patches_forward(_ptr(xp), _ptr(wt), _ptr(bias), _ptr(out), n, hp - 2, wp - 2, kp, cout, int(relu),
                                             int(unpadded), _ptr(ws), ws.numel(), _stream()), "azn_conv_patches_forward")
    return out


def maxpool2x2(x: torch.Tensor, out: torch.Tensor | None = None):
    """MAX 2x2 / stride 2 pooling, ceil mode (pooling_layer.cpp:81-95), zero-bordered NHWC bf16 in and out."""
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    n, hp, wp, c = x.shape
    h, w = hp - 2, wp - 2
    shape = (n, (h + 1) // 2 + 2, (w + 1) // 2 + 2, c)
    if out is None:
        out = torch.empty(shape, dtype=torch.bfloat16, device=x.device)
    assert tuple(out.shape) == shape and out.is_contiguous()
    L.check(L.lib().azn_maxpool2x2_forward(_ptr(x), n, h, w, c, _ptr(out), _stream()), "azn_maxpool2x2_forward")
    return out


def nhwc_border(x: torch.Tensor, to_padded: bool, out: torch.Tensor | None = None):
    """[n, H, W, C] bf16 -> zero-bordered [n, H+2, W+2, C] (to_padded) or back."""
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    n, a, b, c = x.shape
    h, w = (a, b) if to_padded else (a - 2, b - 2)
    shape = (n, h + 2, w + 2, c) if to_padded else (n, h, w, c)
    if out is None:
        out = torch.empty(shape, dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().azn_nhwc_border(_ptr(x), n, h, w, c, _ptr(out), int(to_padded), _stream()), "azn_nhwc_border")
    return out
