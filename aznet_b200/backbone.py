"""VGG16 conv1_1 .. conv5_3 (models/Pascal/VGG16/az-net/test.prototxt:16-384) producing the shared
conv5_3 map.  This is NOT part of the hand-written hot path ("next" row): it runs once per
image before the search and is served by PyTorch/cuDNN as plumbing so that the from-image entry points
(`im_propose(net, im)`) are usable.  Pooling is ceil-mode like Caffe (pooling_layer.cpp:93-95)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

VGG16_CFG = [(64, 2), (128, 2), (256, 3), (512, 3), (512, 3)]      # (channels, convs) per stage


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


class VGG16Backbone:
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
