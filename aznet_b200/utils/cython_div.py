"""Drop-in for utils.cython_div (lib/utils/div.pyx:15-88): divide_region(regions, min_height) and
_sift_dup(regions, min_height) on float64 [N,4] ndarrays, computed by libaznet_b200.so."""
import numpy as np
import torch

from .. import ops


def _run(regions, min_height, sift_only):
    if not isinstance(regions, np.ndarray) or regions.dtype != np.float64 or regions.ndim != 2:
        raise ValueError("Buffer dtype mismatch, expected 'float_t' 2-D array")     # Cython typed buffer, div.pyx:15
    if regions.shape[0] == 0:
        return np.zeros((0, 4), dtype=np.float64)
    r = torch.from_numpy(np.ascontiguousarray(regions[:, :4])).cuda()
    out, cnt = ops.divide_region(r, float(min_height), sift_only=sift_only)
    n = int(cnt.item())
    if n < 0:
        raise RuntimeError("divide_region: scratch capacity exceeded (degenerate aspect ratio?)")
    return out[:n].cpu().numpy()


def divide_region(regions, min_height):
    return _run(regions, min_height, False)


def _sift_dup(regions, min_height):
    return _run(regions, min_height, True)
