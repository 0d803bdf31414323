"""Duck-typed stand-in for the `caffe.Net` objects that lib/detect/test.py programs against
(caffe-fast-rcnn/python/caffe/pycaffe.py:21-95, _caffe.cpp:75-244): `.blobs[name].reshape(*shape)`,
`.forward(blobs=[...], **inputs) -> {name: ndarray}`, `.inputs`, `.outputs`, `.name`.

`Net(kind='az')` is the AZ-Net of models/*/VGG16/az-net/test{,_fc}.prototxt, `Net(kind='frcnn')` the Fast
R-CNN detector of models/*/VGG16/frcnn/test{,_fc}.prototxt.  With a `backbone` the net is the 'full' one
(inputs data + rois); without, the 'fc' one (inputs conv5_3 + rois).  Every forward runs on the GPU through
the C ABI: ROI max-pool -> tcgen05 fc layers.  There is no CPU path.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

from . import _lib as L
from . import ops
from .engine import AZHeadWeights


class FRCNNHeadWeights:
    """Fast R-CNN fc weights prepared for the tensor-core path: fc6 (K permuted), fc7, and cls_score |
    bbox_pred fused into one N = 5*C GEMM whose first C columns go through the softmax."""

    def __init__(self, weights: dict, device, pooled=7):
        g = lambda n: (torch.from_numpy(np.ascontiguousarray(weights[n][0])), torch.from_numpy(np.ascontiguousarray(weights[n][1])))
        w6, b6 = g("fc6")
        self.h6, k6 = w6.shape
        self.pooled, self.C = pooled, k6 // (pooled * pooled)
        w6 = w6.to(device).view(self.h6, self.C, pooled * pooled).transpose(1, 2).contiguous().view(self.h6, k6)
        self.w6, self.b6 = w6.to(torch.bfloat16).contiguous(), b6.to(device).float().contiguous()
        w7, b7 = g("fc7")
        self.w7, self.b7 = w7.to(device).to(torch.bfloat16).contiguous(), b7.to(device).float().contiguous()
        wc, bc = g("cls_score")
        wb, bb = g("bbox_pred")
        self.num_classes = wc.shape[0]
        self.wo = torch.cat([wc, wb], 0).to(device).to(torch.bfloat16).contiguous()
        self.bo = torch.cat([bc, bb], 0).to(device).float().contiguous()


class FRCNNSkipHeadWeights(FRCNNHeadWeights):
    """The skip-layer detector head of models/COCO/VGG16_skip/frcnn/test_fc.prototxt:28-232: ROI pools
    over conv3_3 / conv4_3 / conv5_3 at 1/4, 1/8, 1/16, GRN each, concat, x1000, conv_pool5 (1x1, + ReLU), then the
    plain Fast R-CNN fc layers over the c_out-channel pool5.  conv_pool5 is a GEMM over pooled positions; its K is
    zero-padded to the tensor-core kernel's K granularity (64)."""
    conv_names = ("conv3_3", "conv4_3", "conv5_3")
    scales = (0.25, 0.125, 0.0625)
    grn_scale = 1000.0

    def __init__(self, weights: dict, device, pooled=7):
        FRCNNHeadWeights.__init__(self, weights, device, pooled)
        wc, bc = weights["conv_pool5"]
        wc = torch.from_numpy(np.ascontiguousarray(wc)).reshape(wc.shape[0], -1)
        ctot = wc.shape[1]
        self.src_channels = tuple(int(c) for c in weights.get("skip_channels", (ctot // 5, 2 * ctot // 5, 2 * ctot // 5)))
        assert sum(self.src_channels) == ctot and wc.shape[0] == self.C, "conv_pool5 shape does not match the pools / fc6"
        self.k_cat = (ctot + 63) // 64 * 64
        wp = torch.zeros((wc.shape[0], self.k_cat), dtype=torch.float32)
        wp[:, :ctot] = wc
        self.wc = wp.to(device).to(torch.bfloat16).contiguous()
        self.bc = torch.from_numpy(np.ascontiguousarray(bc)).to(device).float().contiguous()


class _Blob:
    def __init__(self):
        self.shape = ()

    def reshape(self, *shape):
        self.shape = tuple(int(s) for s in shape)

    @property
    def num(self):
        return self.shape[0] if self.shape else 0


class Net:
    def __init__(self, weights: dict, kind: str = "az", backbone=None, name: str = "aznet_b200", device=None,
                 pooled: int = 7, spatial_scale: float = 0.0625):
        L.require_device()
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.kind, self.name, self.backbone = kind, name, backbone
        self.pooled, self.spatial_scale = pooled, spatial_scale
        if kind == "az":
            self.head = weights if isinstance(weights, AZHeadWeights) else AZHeadWeights(weights, self.dev, pooled)
            self.outputs = ["zoom_prob", "adj_prob", "adj_bbox"]
        elif kind == "frcnn":
            self.head = weights if isinstance(weights, FRCNNHeadWeights) else FRCNNHeadWeights(weights, self.dev, pooled)
            self.outputs = ["cls_prob", "bbox_pred"]
        elif kind == "frcnn_skip":
            self.head = weights if isinstance(weights, FRCNNSkipHeadWeights) else FRCNNSkipHeadWeights(weights, self.dev, pooled)
            self.outputs = ["cls_prob", "bbox_pred"]
        else:
            raise ValueError("kind must be 'az', 'frcnn' or 'frcnn_skip'")
        self.conv_names = tuple(self.head.conv_names) if kind == "frcnn_skip" else ("conv5_3",)
        self.inputs = ["data", "rois"] if backbone is not None else list(self.conv_names) + ["rois"]
        self.blobs = {k: _Blob() for k in ("data", "rois", "conv5_3") + self.conv_names}
        self._conv_key, self._conv_nhwc, self._conv_dev = None, None, None
        self._maps = {}                         # skip head: name -> (key, f32 NCHW device copy, bf16 NHWC)

    # conv maps handed in as host arrays are uploaded once per distinct array (the reference re-uploads
    # the 4.9 MB map on every call, pycaffe.py:90)
    @staticmethod
    def _map_key(host: np.ndarray):
        """Identity of a host map: object, address, shape AND a content fingerprint (CRC of a strided sample of up to
        8192 elements), so that a caller that refills the same ndarray in place -- pycaffe-style blob views over reused
        net memory do -- gets a fresh upload instead of the previous image's device copy."""
        flat = host.reshape(-1)
        step = max(1, flat.size // 8192)
        return (id(host), host.__array_interface__["data"][0], host.shape, zlib.crc32(np.ascontiguousarray(flat[::step]).tobytes()))

    def _resident_conv(self, conv_host: np.ndarray):
        key = self._map_key(conv_host)
        if key != self._conv_key:
            self._conv_dev = torch.from_numpy(np.ascontiguousarray(conv_host, dtype=np.float32)).to(self.dev)
            self._conv_nhwc = ops.nchw_to_nhwc_bf16(self._conv_dev)
            self._conv_key, self._conv_host = key, conv_host
        return self._conv_nhwc

    def conv_from_data(self, data_host: np.ndarray):
        conv = self.backbone(torch.from_numpy(np.ascontiguousarray(data_host, dtype=np.float32)).to(self.dev))
        return conv, ops.nchw_to_nhwc_bf16(conv)

    def _resident_map(self, name, host: np.ndarray):
        key = self._map_key(host)
        ent = self._maps.get(name)
        if ent is None or ent[0] != key:
            dev = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32)).to(self.dev)
            ent = (key, dev, ops.nchw_to_nhwc_bf16(dev), host)
            self._maps[name] = ent
        return ent[2]

    def skip_pool5(self, maps: dict, rois: torch.Tensor, n_rois: torch.Tensor | None = None, fused: bool = True):
        """The skip head up to pool5: three ROI pools -> GRN + concat + x1000 -> conv_pool5 (+ReLU).
        maps: name -> bf16 NHWC; returns bf16 [R, 49 * c_out] (the pooled-row matrix of fc6)."""
        hd, P = self.head, self.pooled
        R = rois.shape[0]
        cat = torch.zeros((R * P * P, hd.k_cat), dtype=torch.bfloat16, device=rois.device)
        if fused:                                  # GRN + concat + scale in the pooling kernel's epilogue
            off = 0
            for n, sc, c in zip(hd.conv_names, hd.scales, hd.src_channels):
                ops.roi_pool_grn(maps[n], rois, cat, off, P, sc, hd.grn_scale, n_rois=n_rois)
                off += c
        else:
            pooled = [ops.roi_pool(maps[n], rois, P, sc, layout="NHWC", n_rois=n_rois).view(R * P * P, -1)
                      for n, sc in zip(hd.conv_names, hd.scales)]
            ops.grn_concat(pooled, hd.grn_scale, n_units=n_rois, rows_per_unit=P * P, out=cat)
        m_live = None if n_rois is None else n_rois * (P * P)
        return ops.fc_forward(cat, hd.wc, hd.bc, L.ACT_RELU, m_live=m_live).view(R, -1)

    def heads_device(self, nhwc, rois: torch.Tensor):
        """ROI pool + fc layers on device tensors; returns the f32 output matrix [R, ld].
        nhwc: the bf16 NHWC conv5_3 map, or for the skip head a dict name -> map."""
        hd = self.head
        if self.kind == "frcnn_skip":
            a = self.skip_pool5(nhwc, rois)
        else:
            pool = ops.roi_pool(nhwc, rois, self.pooled, self.spatial_scale, layout="NHWC")
            a = pool.view(pool.shape[0], -1)
        if self.kind == "az":
            h6 = ops.fc_forward(a, hd.w6, hd.b6, L.ACT_RELU)
            h7 = ops.fc_forward(h6, hd.w7, hd.b7, L.ACT_RELU)
            out = torch.empty((a.shape[0], hd.ld_head), dtype=torch.float32, device=a.device)
            return ops.az_heads(h7, hd.wh, hd.bh, hd.nsub, out=out)
        h6 = ops.fc_forward(a, hd.w6, hd.b6, L.ACT_RELU)
        h7 = ops.fc_forward(h6, hd.w7, hd.b7, L.ACT_RELU)
        n = hd.wo.shape[0]
        out = torch.empty((a.shape[0], (n + 3) // 4 * 4), dtype=torch.float32, device=a.device)
        return ops.fc_forward(h7, hd.wo, hd.bo, L.ACT_SOFTMAX_BBOX, hd.num_classes, out=out[:, :n])

    def forward(self, blobs=None, start=None, end=None, **kwargs):
        if start is not None or end is not None:
            raise NotImplementedError("partial forward (start/end) is not part of the hot path")
        if set(kwargs.keys()) != set(self.inputs):
            raise Exception("Input blob arguments do not match net inputs.")          # pycaffe.py:83-84
        for k, v in kwargs.items():
            if v.shape[0] != self.blobs[k].num:
                raise Exception("Input is not batch sized")                          # pycaffe.py:88-89
        rois_h = np.ascontiguousarray(kwargs["rois"], dtype=np.float32)
        out, conv_host = {}, None
        skip_hosts = None
        if self.kind == "frcnn_skip":
            if self.backbone is not None:
                data = torch.from_numpy(np.ascontiguousarray(kwargs["data"], dtype=np.float32)).to(self.dev)
                nhwc = self.backbone.taps_from_data(data, self.conv_names)
            else:
                nhwc = {n: self._resident_map(n, kwargs[n]) for n in self.conv_names}
                skip_hosts = {n: kwargs[n] for n in self.conv_names}
            conv_dev = None
        elif self.backbone is not None:
            extra = [b for b in (blobs or []) if b != "conv5_3" and b in getattr(self.backbone, "names", ())]
            if extra:
                # SEAR.FRCNN_CONV of a skip-layer config (voc_skip.yml:20) asked of the AZ 'full' net (test.py:222-226):
                # the other taps of the same backbone pass are handed out as Caffe-style blobs below
                data = torch.from_numpy(np.ascontiguousarray(kwargs["data"], dtype=np.float32)).to(self.dev)
                taps = self.backbone.taps_from_data(data, tuple(dict.fromkeys(extra + ["conv5_3"])))
                nhwc = taps["conv5_3"]
                conv_dev = nhwc.permute(0, 3, 1, 2).float().contiguous()
                skip_hosts = {}
                for b in extra:
                    dev = taps[b].permute(0, 3, 1, 2).float().contiguous()
                    skip_hosts[b] = dev.cpu().numpy()
                    self._maps[b] = (self._map_key(skip_hosts[b]), dev, taps[b], skip_hosts[b])
            else:
                conv_dev, nhwc = self.conv_from_data(kwargs["data"])
        else:
            conv_host = kwargs["conv5_3"]
            nhwc = self._resident_conv(conv_host)
            conv_dev = self._conv_dev
        R = rois_h.shape[0]
        if R > 0:
            res = self.heads_device(nhwc, torch.from_numpy(rois_h).to(self.dev)).cpu().numpy()
        else:
            res = np.zeros((0, 5 * getattr(self.head, "nsub", 0) + 1 if self.kind == "az" else 5 * self.head.num_classes), np.float32)
        if self.kind == "az":
            ns = self.head.nsub
            out["adj_prob"] = np.ascontiguousarray(res[:, :ns])
            out["adj_bbox"] = np.ascontiguousarray(res[:, ns:5 * ns])
            out["zoom_prob"] = np.ascontiguousarray(res[:, 5 * ns:5 * ns + 1])
        else:
            c = self.head.num_classes
            out["cls_prob"] = np.ascontiguousarray(res[:, :c])
            out["bbox_pred"] = np.ascontiguousarray(res[:, c:5 * c])
        for b in (blobs or []):
            if self.kind == "frcnn_skip" and b in self.conv_names:
                if skip_hosts is None:                     # maps computed by the backbone: hand them out as Caffe blobs
                    skip_hosts = {}
                if b not in skip_hosts:
                    dev = nhwc[b].permute(0, 3, 1, 2).float().contiguous()
                    skip_hosts[b] = dev.cpu().numpy()
                    self._maps[b] = (self._map_key(skip_hosts[b]), dev, nhwc[b], skip_hosts[b])
                out[b] = skip_hosts[b]
            elif skip_hosts is not None and b in skip_hosts:
                out[b] = skip_hosts[b]
            elif b == "conv5_3":
                if conv_host is None:
                    conv_host = conv_dev.cpu().numpy()
                    # remember the upload so that handing this array back to the 'fc' net costs nothing
                    self._last_conv = (conv_host, conv_dev, nhwc)
                out[b] = conv_host
            else:
                raise KeyError("blob %r is not exposed by the B200 net" % b)
        return out

    def adopt_conv(self, other: "Net"):
        """Share the device copy of a conv map produced by another net's forward (full -> fc hand-over)."""
        for name, ent in getattr(other, "_maps", {}).items():
            if name in self.conv_names and self.kind == "frcnn_skip":
                self._maps[name] = ent
        last = getattr(other, "_last_conv", None)
        if last is not None:
            host, dev, nhwc = last
            self._conv_key = self._map_key(host)
            self._conv_dev, self._conv_nhwc, self._conv_host = dev, nhwc, host
