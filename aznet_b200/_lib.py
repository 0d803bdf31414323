"""ctypes binding of libaznet_b200.so (include/aznet_b200.h).

The shared library is built in-tree by `build()` (nvcc, sm_100a only) and is the ONLY compute
path: there is no CPU or PyTorch fallback.  `lib()` raises if the library is missing.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "libaznet_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "aznet_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",           # one rounding per operation: NMS / region arithmetic must be bit-exact
              "-Xcompiler", "-fPIC", "-shared"]

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY = 0, 1, 2, 3
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
DTYPE_F32, DTYPE_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_AZ_HEAD, ACT_SOFTMAX_BBOX = 0, 1, 2, 3
NMS_SEG_MAX = 1024
LEVEL_LAST, LEVEL_ROOT_PROPS, LEVEL_TUNE = 1, 2, 4

EXPORTS = [
    "azn_version", "azn_last_error", "azn_check_device", "azn_set_pdl", "azn_set_coop", "azn_hbm_write_probe", "azn_roi_pool_workspace_bytes", "azn_roi_pool_tune", "azn_roi_pool_fwd", "azn_roi_pool_fwd_ex", "azn_nchw_f32_to_nhwc_bf16", "azn_nchw_bf16_to_nhwc_bf16", "azn_host_f32_to_bf16", "azn_host_threads",
    "azn_fc_workspace_bytes", "azn_fc_tune", "azn_fc_trace", "azn_fc_forward", "azn_az_heads_forward", "azn_az_heads_tune", "azn_search_init", "azn_search_root", "azn_search_level", "azn_select_proposals", "azn_collect_proposals",
    "azn_peer_alloc", "azn_peer_open", "azn_peer_close", "azn_peer_free",
    "azn_divide_region", "azn_divide_region_scratch_bytes", "azn_decode_boxes", "azn_nms_workspace_bytes",
    "azn_nms", "azn_nms_batched", "azn_nms_segments", "azn_nms_tune",
    "azn_detect_rois", "azn_detect_select", "azn_detect_thresholds", "azn_detect_filter", "azn_tune_threshold",
    "azn_image_blob", "azn_conv3x3_forward", "azn_conv_tune", "azn_maxpool2x2_forward", "azn_nhwc_border", "azn_grn_concat_forward", "azn_roi_pool_grn_fwd", "azn_patches3x3", "azn_conv_patches_forward", "azn_conv3x3_direct_forward",
]


def sources():
    """CUDA sources (nvcc) and the host-only C++ sources (g++: no device code, x86 intrinsics)."""
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu"))) + sorted(glob.glob(os.path.join(_CSRC, "*.cpp")))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into aznet_b200/libaznet_b200.so (nvcc cross-compiles
    without a GPU)."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(_CSRC, "*.cuh")) + [HEADER]
    if not force and os.path.exists(SO_PATH) and all(os.path.getmtime(SO_PATH) >= os.path.getmtime(d) for d in deps):
        return SO_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    build_dir = os.path.join(_HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for s in srcs:
        o = os.path.join(build_dir, os.path.splitext(os.path.basename(s))[0] + ".o")
        objs.append(o)
        if s.endswith(".cpp"):
            cmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-pthread", "-c", s, "-o", o]
        else:
            cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("AZN_NVCC_EXTRA", "").split() + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, out))
        if verbose and out:
            print(out)
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-pthread", "-o", SO_PATH] + objs)
    return SO_PATH


class SearchState(C.Structure):
    """Mirror of `struct azn_search_state`."""
    _fields_ = [
        ("n_img", C.c_int32), ("cap_regions", C.c_int32), ("cap_children", C.c_int32), ("cap_props", C.c_int32),
        ("nsub", C.c_int32), ("chunk", C.c_int32),
        ("im_h", C.c_void_p), ("im_w", C.c_void_p), ("im_scale", C.c_void_p),
        ("tz", C.c_double), ("min_side", C.c_double), ("eps", C.c_double), ("dedup", C.c_double),
        ("regions", C.c_void_p), ("n_regions", C.c_void_p), ("inv", C.c_void_p), ("rep", C.c_void_p),
        ("n_uniq", C.c_void_p), ("img_off", C.c_void_p), ("rois", C.c_void_p), ("m_total", C.c_void_p),
        ("next_regions", C.c_void_p), ("next_n_regions", C.c_void_p),
        ("children", C.c_void_p), ("hashes", C.c_void_p), ("flags", C.c_void_p),
        ("props", C.c_void_p), ("prop_scores", C.c_void_p), ("n_props", C.c_void_p),
        ("n_eval", C.c_void_p), ("depth", C.c_void_p), ("status", C.c_void_p),
        ("hist_regions", C.c_void_p), ("hist_zoom", C.c_void_p), ("n_history", C.c_void_p), ("cap_history", C.c_int32),
    ]


class DetectState(C.Structure):
    """Mirror of `struct azn_detect_state`."""
    _fields_ = [
        ("n_img", C.c_int32), ("cap_boxes", C.c_int32), ("num_classes", C.c_int32), ("max_per_image", C.c_int32),
        ("chunk", C.c_int32), ("ld_head", C.c_int32),
        ("im_h", C.c_void_p), ("im_w", C.c_void_p), ("im_scale", C.c_void_p),
        ("eps", C.c_double), ("dedup", C.c_double),
        ("boxes", C.c_void_p), ("n_boxes", C.c_void_p), ("inv", C.c_void_p), ("rep", C.c_void_p),
        ("n_uniq", C.c_void_p), ("img_off", C.c_void_p), ("rois", C.c_void_p), ("m_total", C.c_void_p),
        ("hashes", C.c_void_p), ("flags", C.c_void_p),
        ("head_out", C.c_void_p), ("thresh", C.c_void_p), ("dets", C.c_void_p), ("top_scores", C.c_void_p),
        ("det_count", C.c_void_p),
    ]


_LIB = None


def _bind(L):
    vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
    L.azn_version.restype = C.c_char_p
    L.azn_last_error.restype = C.c_char_p
    L.azn_check_device.restype = i32
    L.azn_roi_pool_fwd.restype = i32
    L.azn_roi_pool_fwd.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, f32, vp, vp, vp, sz, vp]
    L.azn_roi_pool_fwd_ex.restype = i32
    L.azn_roi_pool_fwd_ex.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, f32, vp, vp, vp, sz, i32, vp]
    L.azn_roi_pool_workspace_bytes.restype = sz
    L.azn_roi_pool_workspace_bytes.argtypes = [i32, i32, i32, i32, i32, i32, i32]
    L.azn_roi_pool_tune.restype = None
    L.azn_roi_pool_tune.argtypes = [i32]
    L.azn_nchw_f32_to_nhwc_bf16.restype = i32
    L.azn_nchw_f32_to_nhwc_bf16.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.azn_nchw_bf16_to_nhwc_bf16.restype = i32
    L.azn_nchw_bf16_to_nhwc_bf16.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.azn_host_f32_to_bf16.restype = i32
    L.azn_host_f32_to_bf16.argtypes = [vp, vp, sz, i32]
    L.azn_host_threads.restype = i32
    L.azn_host_threads.argtypes = []
    L.azn_fc_workspace_bytes.restype = sz
    L.azn_fc_workspace_bytes.argtypes = [i32, i32, i32]
    L.azn_fc_tune.restype = None
    L.azn_fc_tune.argtypes = [i32, i32, i32]
    L.azn_set_pdl.restype = None
    L.azn_set_pdl.argtypes = [i32]
    L.azn_set_coop.restype = None
    L.azn_set_coop.argtypes = [i32]
    L.azn_hbm_write_probe.restype = i32
    L.azn_hbm_write_probe.argtypes = [vp, sz, vp]
    L.azn_fc_trace.restype = None
    L.azn_fc_trace.argtypes = [vp]
    L.azn_fc_forward.restype = i32
    L.azn_fc_forward.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, i32, i32, i32, i32, vp, sz, vp]
    L.azn_az_heads_tune.restype = None
    L.azn_az_heads_tune.argtypes = [i32]
    L.azn_az_heads_forward.restype = i32
    L.azn_az_heads_forward.argtypes = [vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, vp]
    L.azn_search_init.restype = i32
    L.azn_search_init.argtypes = [C.POINTER(SearchState), vp]
    L.azn_search_root.restype = i32
    L.azn_search_root.argtypes = [C.POINTER(SearchState), vp]
    L.azn_search_level.restype = i32
    L.azn_search_level.argtypes = [C.POINTER(SearchState), vp, i32, vp, i32, vp, i32, i32, i32, vp]
    L.azn_select_proposals.restype = i32
    L.azn_select_proposals.argtypes = [C.POINTER(SearchState), i32, i32, f64, vp, vp, vp, i32, vp]
    L.azn_collect_proposals.restype = i32
    L.azn_collect_proposals.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, i32, vp, vp]
    L.azn_conv_tune.restype = None
    L.azn_conv_tune.argtypes = [i32, i32]
    L.azn_peer_alloc.restype = i32
    L.azn_peer_alloc.argtypes = [sz, C.POINTER(C.c_void_p), C.c_char_p]
    L.azn_peer_open.restype = i32
    L.azn_peer_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.azn_peer_close.restype = i32
    L.azn_peer_close.argtypes = [vp]
    L.azn_peer_free.restype = i32
    L.azn_peer_free.argtypes = [vp]
    L.azn_divide_region.restype = i32
    L.azn_divide_region.argtypes = [vp, i32, f64, vp, vp, i32, i32, vp, sz, vp]
    L.azn_divide_region_scratch_bytes.restype = sz
    L.azn_divide_region_scratch_bytes.argtypes = [i32]
    L.azn_decode_boxes.restype = i32
    L.azn_decode_boxes.argtypes = [vp, vp, i32, i32, f64, i32, i32, vp, vp]
    L.azn_nms_workspace_bytes.restype = sz
    L.azn_nms_workspace_bytes.argtypes = [i64]
    L.azn_nms.restype = i32
    L.azn_nms.argtypes = [vp, i64, f64, vp, vp, vp, sz, vp]
    L.azn_nms_batched.restype = i32
    L.azn_nms_batched.argtypes = [vp, vp, i32, f64, vp, vp, vp]
    L.azn_nms_segments.restype = i32
    L.azn_nms_segments.argtypes = [vp, vp, vp, i32, i32, f64, vp, vp, vp]
    L.azn_detect_rois.restype = i32
    L.azn_detect_rois.argtypes = [C.POINTER(DetectState), vp]
    L.azn_detect_select.restype = i32
    L.azn_detect_select.argtypes = [C.POINTER(DetectState), vp]
    L.azn_detect_thresholds.restype = i32
    L.azn_detect_thresholds.argtypes = [vp, vp, i32, i32, i32, C.c_longlong, vp, vp]
    L.azn_nms_tune.restype = None
    L.azn_nms_tune.argtypes = [i32]
    L.azn_roi_pool_grn_fwd.restype = i32
    L.azn_roi_pool_grn_fwd.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, i32, i32, f32, f32, vp, i32, i32, vp]
    L.azn_grn_concat_forward.restype = i32
    L.azn_grn_concat_forward.argtypes = [vp, vp, i32, vp, C.c_longlong, i32, f32, vp, i32, vp]
    L.azn_tune_threshold.restype = i32
    L.azn_tune_threshold.argtypes = [vp, vp, i32, i32, C.c_longlong, vp, vp]
    L.azn_detect_filter.restype = i32
    L.azn_detect_filter.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.azn_image_blob.restype = i32
    L.azn_image_blob.argtypes = [vp, i32, i32, i32, f64, vp, vp, i32, i32, i32, vp, vp]
    L.azn_conv3x3_forward.restype = i32
    L.azn_conv3x3_forward.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, sz, vp]
    L.azn_conv_patches_forward.restype = i32
    L.azn_conv_patches_forward.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, sz, vp]
    L.azn_patches3x3.restype = i32
    L.azn_patches3x3.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32, vp]
    L.azn_conv3x3_direct_forward.restype = i32
    L.azn_conv3x3_direct_forward.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, i32, i32, vp]
    L.azn_maxpool2x2_forward.restype = i32
    L.azn_maxpool2x2_forward.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.azn_nhwc_border.restype = i32
    L.azn_nhwc_border.argtypes = [vp, i32, i32, i32, i32, vp, i32, vp]
    return L


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("libaznet_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for the AZ-Net hot path)")
        _LIB = _bind(C.CDLL(SO_PATH))
    return _LIB


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().azn_last_error().decode()
        kind = {1: "invalid argument", 2: "CUDA error", 3: "capacity"}.get(rc, "error %d" % rc)
        if rc == ERR_INVALID:
            raise ValueError("%s: %s (%s)" % (what or "aznet_b200", msg, kind))
        raise RuntimeError("%s: %s (%s)" % (what or "aznet_b200", msg, kind))


def require_device():
    """Raise unless a B200-class (sm_100) CUDA device is current."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("aznet_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
    check(lib().azn_check_device(), "azn_check_device")
