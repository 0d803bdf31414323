"""Drop-in for utils.cython_nms (lib/utils/nms.pyx:17-68): nms(dets, thresh) -> list of kept indices,
computed by the bitmask NMS kernels of libaznet_b200.so.  Same argument contract as the Cython module:
`dets` must be a 2-D float32 ndarray (anything else -> ValueError, like the typed buffer) and `thresh`
a Python float (TypeError otherwise)."""
import numpy as np
import torch

from .. import ops


def nms(dets, thresh):
    if not isinstance(dets, np.ndarray) or dets.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got %r" % (getattr(dets, "dtype", type(dets)),))
    if dets.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % dets.ndim)
    if not isinstance(thresh, float):
        raise TypeError("Argument 'thresh' has incorrect type (expected float, got %s)" % type(thresh).__name__)
    if dets.shape[0] == 0:
        return []
    if dets.shape[1] < 5:
        raise IndexError("dets needs 5 columns (x1, y1, x2, y2, score)")
    d = torch.from_numpy(np.ascontiguousarray(dets[:, :5])).cuda(non_blocking=True)
    keep, cnt = ops.nms(d, thresh)
    return keep[:int(cnt.item())].cpu().tolist()
