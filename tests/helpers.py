"""Shared helpers of the GPU drop-in tests."""
import numpy as np
import torch


def _same_blob_for_both_routes(monkeypatch, dev):
    """The fast route builds the network input on the device (azn_image_blob, bf16); the host route calls
    T._get_image_blob (cv2).  The two blobs agree to 4 float32 ulps (tests/test_gpu_backbone.py), which is enough
    to flip a bf16 rounding here and there.  To compare the ROUTES and not the resizers, hand the host route the
    product's own blob (the f32 'data' blob of the same kernel, rounded to bf16 like the backbone's input)."""
    from aznet_b200 import backbone, ops
    from aznet_b200.detect import test as T
    from aznet_b200.engine import im_scale_for

    def blob(im):
        s = im_scale_for(im.shape[0], im.shape[1], tuple(T.cfg.TEST.SCALES), T.cfg.TEST.MAX_SIZE)
        pix = torch.from_numpy(np.ascontiguousarray(im)[None]).to(dev)
        _, f32 = ops.image_blob(pix, s, backbone.PIXEL_MEANS, 8, want_f32=True)
        return f32.to(torch.bfloat16).float().cpu().numpy(), np.array([s])
    monkeypatch.setattr(T, "_get_image_blob", blob)


class RowMatcher:
    """Accumulates how many rows of paired row sets have a partner within `atol` in the other set.  Batched and
    per-image passes run the same kernels but split the GEMM reductions differently, so a few bf16 activations round
    the other way (scores move by ~1e-3, a borderline row may enter or leave a selection): the routes must agree on
    >= `frac` of all rows, not on every one."""

    def __init__(self, atol):
        self.atol, self.hit, self.total = atol, 0, 0

    def add(self, a, b):
        a = np.zeros((0, 1)) if isinstance(a, list) else np.asarray(a, dtype=np.float64)
        b = np.zeros((0, 1)) if isinstance(b, list) else np.asarray(b, dtype=np.float64)
        self.total += len(a) + len(b)
        if len(a) and len(b):
            d = np.abs(a[:, None, :] - b[None, :, :]).max(2)
            self.hit += int((d.min(1) <= self.atol).sum() + (d.min(0) <= self.atol).sum())

    def check(self, frac, min_total=1):
        assert self.total >= min_total, self.total
        assert self.hit >= frac * self.total, (self.hit, self.total)


def _assert_rows_match(a, b, atol, frac=0.98):
    m = RowMatcher(atol)
    m.add(a, b)
    m.check(frac, 0)
