"""oracle/ref_caffe.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings of oracle/_ref/libcaffe_layers_ref.so: the reference's OWN layer sources
(caffe-fast-rcnn/src/caffe/layers/{roi_pooling,grn,sigmoid,softmax,relu,inner_product,pooling}_layer.cpp,
unmodified, read from /root/reference) compiled by oracle/build_ref.py against the stand-in framework
headers of oracle/caffe_shim.  Used to pin the restatements in oracle/caffe_layers.c / az_oracle.py
(tests/test_oracle.py) and to generate tests/golden/caffe_layers.npz (oracle/gen_golden.py).
The library is prebuilt in the authoring container and travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libcaffe_layers_ref.so")
_LIB = None


def available() -> bool:
    return os.path.exists(SO)


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(SO)
        c = ctypes
        p, i, f = c.c_void_p, c.c_int, c.c_float
        for name, args in (("ref_roi_pool_fwd", [p, i, i, i, i, p, i, i, i, f, p, p]),
                           ("ref_grn_fwd", [p, i, i, i, i, p]),
                           ("ref_sigmoid_fwd", [p, p, i]),
                           ("ref_relu_fwd", [p, p, i]),
                           ("ref_softmax_fwd", [p, p, i, i]),
                           ("ref_inner_product_fwd", [p, i, i, p, p, i, p]),
                           ("ref_max_pool_fwd", [p, i, i, i, i, i, i, i, p, i, p, p])):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = c.c_int, args
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def roi_pool_fwd(feat, rois, pooled=7, spatial_scale=0.0625, want_argmax=False):
    """ROIPoolingLayer<float>::Forward_cpu, roi_pooling_layer.cpp:46-125."""
    feat, rois = _f32(feat), _f32(rois).reshape(-1, 5)
    n, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.empty((R, C, pooled, pooled), np.float32)
    amax = np.empty((R, C, pooled, pooled), np.int32) if want_argmax else None
    rc = _lib().ref_roi_pool_fwd(feat.ctypes.data, n, C, H, W, rois.ctypes.data, R, pooled, pooled,
                                 ctypes.c_float(spatial_scale), out.ctypes.data, amax.ctypes.data if want_argmax else None)
    if rc != 0:
        raise RuntimeError("CHECK failed in ROIPoolingLayer::Forward_cpu (the reference aborts here)")
    return (out, amax) if want_argmax else out


def grn(x):
    """GRNLayer<float>::Forward_cpu, grn_layer.cpp:27-56."""
    x = _f32(x)
    y = np.empty_like(x)
    assert _lib().ref_grn_fwd(x.ctypes.data, *x.shape, y.ctypes.data) == 0
    return y


def sigmoid(x):
    """SigmoidLayer<float>::Forward_cpu, sigmoid_layer.cpp:11-24."""
    x = _f32(x)
    y = np.empty_like(x)
    assert _lib().ref_sigmoid_fwd(x.ctypes.data, y.ctypes.data, x.size) == 0
    return y


def relu(x):
    """ReLULayer<float>::Forward_cpu, relu_layer.cpp:10-20."""
    x = _f32(x)
    y = np.empty_like(x)
    assert _lib().ref_relu_fwd(x.ctypes.data, y.ctypes.data, x.size) == 0
    return y


def softmax(x):
    """SoftmaxLayer<float>::Forward_cpu, softmax_layer.cpp:28-60, rows of [R, C]."""
    x = _f32(x)
    y = np.empty_like(x)
    assert _lib().ref_softmax_fwd(x.ctypes.data, y.ctypes.data, x.shape[0], x.shape[1]) == 0
    return y


def inner_product(x, w, b):
    """InnerProductLayer<float>::Forward_cpu, inner_product_layer.cpp:80-93 (BLAS calls resolved by the shim's
    netlib-order gemm: scalar, so keep M*N*K small)."""
    x, w = _f32(x), _f32(w)
    x2 = x.reshape(x.shape[0], -1)
    b = None if b is None else _f32(b)
    y = np.empty((x2.shape[0], w.shape[0]), np.float32)
    rc = _lib().ref_inner_product_fwd(x2.ctypes.data, x2.shape[0], x2.shape[1], w.ctypes.data,
                                      None if b is None else b.ctypes.data, w.shape[0], y.ctypes.data)
    if rc != 0:
        raise RuntimeError("CHECK failed in InnerProductLayer (input size incompatible)")
    return y


def max_pool(x, kernel=2, stride=2, pad=0):
    """PoolingLayer<float> (MAX): Reshape :84-126 (ceil mode) + Forward_cpu :128-229."""
    x = _f32(x)
    n, C, H, W = x.shape
    cap = n * C * (H // stride + 2) * (W // stride + 2)
    y = np.empty(cap, np.float32)
    ph, pw = ctypes.c_int(), ctypes.c_int()
    rc = _lib().ref_max_pool_fwd(x.ctypes.data, n, C, H, W, kernel, stride, pad, y.ctypes.data, cap,
                                 ctypes.byref(ph), ctypes.byref(pw))
    assert rc == 0, rc
    return y[:n * C * ph.value * pw.value].reshape(n, C, ph.value, pw.value).copy()
