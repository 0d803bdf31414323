"""Image -> network blob helpers with the semantics of lib/utils/blob.py:13-45 (host-side boundary code)."""
import cv2
import numpy as np


def im_list_to_blob(ims):
    """List of HWC float images -> zero-padded NCHW float32 blob (blob.py:13-29)."""
    shape = np.array([im.shape for im in ims]).max(axis=0)
    blob = np.zeros((len(ims), shape[0], shape[1], shape[2]), dtype=np.float32)
    for i, im in enumerate(ims):
        blob[i, :im.shape[0], :im.shape[1], :] = im
    return blob.transpose((0, 3, 1, 2))


def prep_im_for_blob(im, pixel_means, target_size, max_size):
    """Mean-subtract and bilinear-rescale one image (blob.py:31-45)."""
    im = im.astype(np.float32, copy=False)
    im -= pixel_means
    size_min, size_max = np.min(im.shape[0:2]), np.max(im.shape[0:2])
    scale = float(target_size) / float(size_min)
    if np.round(scale * size_max) > max_size:
        scale = float(max_size) / float(size_max)
    im = cv2.resize(im, None, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
    return im, scale
