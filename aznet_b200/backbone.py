"""VGG16 conv1_1 .. conv5_3 (models/Pascal/VGG16/az-net/test.prototxt:16-384) producing the shared conv5_3
map -- SURVEY 8f-1, the first "next" row after the search path itself.

`VGG16Native` is the product: image blob (azn_image_blob), thirteen 3x3 convolutions as implicit GEMMs on the
tcgen05 tensor cores (azn_conv3x3_forward: the fc GEMM kernel fed with row-shifted TMA boxes over zero-bordered
channels-last maps) and four ceil-mode 2x2 max pools (azn_maxpool2x2_forward), all through the C ABI.  Pooling is
ceil-mode like Caffe (pooling_layer.cpp:93-95): 600x1000 -> 38x63.

`VGG16Torch` is the same network through PyTorch/cuDNN.  It is a LIBRARY reference for tests and for the
"vs cuDNN" line of tools/backbone_bench.py only; nothing in the product path constructs it.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

VGG16_CFG = [(64, 2), (128, 2), (256, 3), (512, 3), (512, 3)]      # (channels, convs) per stage
PIXEL_MEANS = (102.9801, 115.9465, 122.7717)                        # lib/detect/config.py:210 (BGR)


def make_vgg16_weights(seed=5, in_ch=3, width_div=1):
    """He-normal conv weights, dict 'convS_I' -> (W [O,I,3,3] f32, b [O] f32).  width_div shrinks the
    channel counts for tests."""
    rng = np.random.default_rng(seed)
    w, c_in = {}, in_ch
    for s, (ch, n) in enumerate(VGG16_CFG, 1):
        ch = max(ch // width_div, 8)
        for i in range(1, n + 1):
            std = np.sqrt(2.0 / (c_in * 9))
            w["conv%d_%d" % (s, i)] = ((rng.standard_normal((ch, c_in, 3, 3), dtype=np.float32) * std).astype(np.float32),
                                       np.zeros((ch,), np.float32))
            c_in = ch
    return w


def _pad64(c):
    return (c + 63) // 64 * 64


class VGG16Native:
    """The backbone on hand-written sm_100a kernels.  Channel counts that are not multiples of 64 (reduced test
    networks, the 3 input channels) are zero-padded in the packed weights: a padded output channel has zero
    weights and zero bias, so it stays exactly 0 through ReLU and pooling."""

    def __init__(self, weights: dict, device, pixel_means=PIXEL_MEANS):
        self.dev = torch.device(device)
        self.pixel_means = tuple(float(m) for m in pixel_means)
        self.layers, self.dims, self.names = [], [], []
        self.first_patches = False
        self.first_direct = False
        c_in = None
        for s, (_, n) in enumerate(VGG16_CFG, 1):
            for i in range(1, n + 1):
                W, b = weights["conv%d_%d" % (s, i)]
                W = torch.from_numpy(np.ascontiguousarray(W)).to(self.dev)
                co, ci = W.shape[0], W.shape[1]
                self.dims.append((ci, co, i == n and s < 5))
                self.names.append("conv%d_%d" % (s, i))
                if c_in is None:
                    self.in_channels = ci
                cop, cip = _pad64(co), _pad64(ci)
                Wp = torch.zeros((cop, ci, 3, 3), dtype=torch.float32, device=self.dev)
                Wp[:co] = W
                bp = torch.zeros((cop,), dtype=torch.float32, device=self.dev)
                bp[:co] = torch.from_numpy(np.ascontiguousarray(b)).to(self.dev)
                # the first layer with few input channels (3 -> 27 real K values): gathered patches, one 64-deep tap
                if not self.layers:
                    self.first_patches = 9 * ci <= 64
                    # ... and when the real K fits 32 and the layer has 64 filters (VGG16's conv1_1) even the patch matrix is
                    # skipped: one mma.sync kernel straight from the network input (AZN_CONV1_PATCHES=1: the patches route, A/B)
                    self.first_direct = 9 * ci <= 32 and cop == 64 and not os.environ.get("AZN_CONV1_PATCHES")     # (the input grid then has 8 channels)
                if not self.layers and self.first_patches:
                    self.layers.append((ops.pack_patch_weight(Wp, 64), bp, i == n and s < 5))
                else:
                    self.layers.append((ops.pack_conv_weight(Wp, cip), bp, i == n and s < 5))
                c_in = co
        self.out_channels = c_in
        # network-input grid: 8 channels per pixel when the first layer runs on patches, else padded to the GEMM's 64
        self.cpad_in = 8 if self.first_patches else _pad64(self.in_channels)
        self.launches_per_call = 1 + len(self.layers) + sum(1 for l in self.layers if l[2]) + int(self.first_patches and not self.first_direct)

    def flops(self, hs: int, ws: int) -> float:
        """Algorithmic FLOPs of conv1_1 .. conv5_3 for one hs x ws network input (real channel counts)."""
        total, h, w = 0.0, hs, ws
        for ci, co, pool in self.dims:
            total += 2.0 * 9 * ci * co * h * w
            if pool:
                h, w = (h + 1) // 2, (w + 1) // 2
        return total

    @torch.no_grad()
    def run_padded(self, x: torch.Tensor, taps=None):
        """x bf16 [n, Hs+2, Ws+2, cpad_in] zero-bordered -> conv5_3 bf16 NHWC [n, fh, fw, C] (post-ReLU).
        With `taps` (layer names, e.g. ('conv3_3', 'conv4_3', 'conv5_3') for the skip-layer detector,
        experiments/cfgs/voc_skip.yml:20) -> dict name -> that layer's post-ReLU map, bf16 NHWC without border."""
        last = len(self.layers) - 1
        out = {}
        for k, (wt, b, pool) in enumerate(self.layers):
            if k == 0 and self.first_patches and self.first_direct and k != last:
                x = ops.conv_direct(x, self.in_channels, wt, b, relu=True)
            elif k == 0 and self.first_patches:
                x = ops.conv_patches(ops.patches3x3(x, self.in_channels, 64), wt, b, relu=True, unpadded=(k == last))
            else:
                x = ops.conv3x3(x, wt, b, relu=True, unpadded=(k == last))
            name = self.names[k]
            if taps and name in taps:
                m = x if k == last else ops.nhwc_border(x, to_padded=False)
                co = self.dims[k][1]
                out[name] = m if m.shape[3] == co else m[..., :co].contiguous()
            if pool:
                x = ops.maxpool2x2(x)
        if taps:
            missing = [t for t in taps if t not in out]
            if missing:
                raise KeyError("unknown backbone layers %s" % missing)
            return out
        if x.shape[3] != self.out_channels:
            x = x[..., :self.out_channels].contiguous()
        return x

    @torch.no_grad()
    def taps_from_data(self, data: torch.Tensor, taps):
        """Caffe's 'data' blob f32 NCHW (device) -> dict layer name -> bf16 NHWC map."""
        n, c, h, w = data.shape
        x = torch.zeros((n, h + 2, w + 2, self.cpad_in), dtype=torch.bfloat16, device=data.device)
        x[:, 1:h + 1, 1:w + 1, :c] = data.permute(0, 2, 3, 1)
        return self.run_padded(x, taps=tuple(taps))

    @torch.no_grad()
    def from_images(self, images: torch.Tensor, im_scale: float) -> torch.Tensor:
        """uint8 [n, H0, W0, 3] BGR images (device) -> conv5_3 bf16 NHWC, everything on the device."""
        return self.run_padded(ops.image_blob(images, im_scale, self.pixel_means, self.cpad_in))

    @torch.no_grad()
    def nhwc_from_data(self, data: torch.Tensor) -> torch.Tensor:
        """Caffe's 'data' blob f32 NCHW [n, 3, H, W] (device) -> conv5_3 bf16 NHWC."""
        n, c, h, w = data.shape
        x = torch.zeros((n, h + 2, w + 2, self.cpad_in), dtype=torch.bfloat16, device=data.device)
        x[:, 1:h + 1, 1:w + 1, :c] = data.permute(0, 2, 3, 1)
        return self.run_padded(x)

    @torch.no_grad()
    def __call__(self, data: torch.Tensor) -> torch.Tensor:
        """The blob-level protocol of Net: 'data' f32 NCHW -> conv5_3 f32 NCHW (both on the device)."""
        return self.nhwc_from_data(data).permute(0, 3, 1, 2).float().contiguous()


class VGG16Torch:
    """Library reference (PyTorch/cuDNN, bf16 channels-last) -- tests and the cuDNN comparison only."""

    def __init__(self, weights: dict, device, dtype=torch.bfloat16):
        self.dev, self.dtype = device, dtype
        self.layers = []
        for s, (_, n) in enumerate(VGG16_CFG, 1):
            for i in range(1, n + 1):
                W, b = weights["conv%d_%d" % (s, i)]
                self.layers.append((torch.from_numpy(W).to(device, dtype).contiguous(memory_format=torch.channels_last),
                                    torch.from_numpy(b).to(device, dtype), i == n and s < 5))
        self.out_channels = self.layers[-1][0].shape[0]

    @torch.no_grad()
    def __call__(self, data: torch.Tensor) -> torch.Tensor:
        """data f32 NCHW [n,3,H,W] (device) -> conv5_3 f32 NCHW [n,C,H/16,W/16] (device, post-ReLU)."""
        x = data.to(self.dtype).contiguous(memory_format=torch.channels_last)
        for W, b, pool in self.layers:
            x = F.relu(F.conv2d(x, W, b, padding=1))
            if pool:
                x = F.max_pool2d(x, 2, 2, ceil_mode=True)
        return x.float().contiguous()


VGG16Backbone = VGG16Native
