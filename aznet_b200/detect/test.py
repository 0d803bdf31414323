"""Drop-in for lib/detect/test.py: the Python entry points that tools/prop_az.py and
tools/test_det_net.py call -- test_proposals, test_net, im_propose, im_detect, im_detect_shared,
apply_nms, divide_region -- with the reference's signatures, argument meaning, printed lines and
pickle formats, computed on the B200 through libaznet_b200.so.

Two execution routes for the adaptive search:
  * nets built from aznet_b200.net.Net: the whole level loop runs device-resident in
    aznet_b200.engine.SearchEngine (no host round trip between levels);
  * any other duck-typed net (e.g. a caffe.Net-like object): the reference's level loop is kept on the
    host, one net.forward per level, while decode/clip, divide_region/_sift_dup and NMS still run in the
    CUDA kernels.
Neither route has a CPU fallback.

The detection drivers (test_net, test_net_shared) likewise: with aznet_b200.net.Net detectors every image's
Fast R-CNN step (ROI dedup, pool, fc head, per-class top-100, decode) runs in aznet_b200.detector.DetectEngine
and stays in HBM; the set-wide thresholds, the final filter and ALL (class, image) NMS problems run once at the
end on the device (the heap cap of the reference is order-independent, see csrc/detect.cu).  Other nets take the
reference's host loop.
"""
from __future__ import annotations

import heapq
import os
import pickle

import numpy as np
import torch

from .config import cfg, get_output_dir
from . import batched
from .. import ops
from ..detector import DetectEngine, DetectionSet
from ..engine import SearchEngine, im_scale_for, search_depth
from ..net import Net
from ..utils import cython_div as div
from ..utils.blob import im_list_to_blob
from ..utils.cython_nms import nms
from ..utils.timer import Timer


# --------------------------------------------------------------------------- blob preparation
def _get_image_blob(im):
    """Mean-subtracted, rescaled image pyramid blob + scale factors (lib/detect/test.py:27-59)."""
    import cv2
    im_orig = im.astype(np.float32, copy=True)
    im_orig -= cfg.PIXEL_MEANS
    size_min, size_max = np.min(im_orig.shape[0:2]), np.max(im_orig.shape[0:2])
    ims, factors = [], []
    for target in cfg.TEST.SCALES:
        s = float(target) / float(size_min)
        if np.round(s * size_max) > cfg.TEST.MAX_SIZE:
            s = float(cfg.TEST.MAX_SIZE) / float(size_max)
        ims.append(cv2.resize(im_orig, None, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR))
        factors.append(s)
    return im_list_to_blob(ims), np.array(factors)


def _project_im_rois(im_rois, scales):
    """lib/detect/test.py:73-97."""
    im_rois = im_rois.astype(np.float64, copy=False)
    if len(scales) > 1:
        w = im_rois[:, 2] - im_rois[:, 0] + 1
        h = im_rois[:, 3] - im_rois[:, 1] + 1
        scaled = (w * h)[:, np.newaxis] * (scales[np.newaxis, :] ** 2)
        levels = np.abs(scaled - 224 * 224).argmin(axis=1)[:, np.newaxis]
    else:
        levels = np.zeros((im_rois.shape[0], 1), dtype=np.int64)
    return im_rois * scales[levels], levels


def _get_rois_blob(im_rois, im_scale_factors):
    """lib/detect/test.py:61-71: [level, x1, y1, x2, y2] float32."""
    rois, levels = _project_im_rois(im_rois, im_scale_factors)
    return np.hstack((levels, rois)).astype(np.float32, copy=False)


def _get_blobs(im, rois):
    blobs = {'data': None, 'rois': None}
    blobs['data'], factors = _get_image_blob(im)
    blobs['rois'] = _get_rois_blob(rois, factors)
    return blobs, factors


# --------------------------------------------------------------------------- box arithmetic (CUDA)
def _bbox_pred_clip(boxes, box_deltas, im_shape):
    """_bbox_pred followed by _clip_boxes (lib/detect/test.py:106-151) in one kernel launch."""
    if boxes.shape[0] == 0:
        return np.zeros((0, box_deltas.shape[1]))
    b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float64)).cuda()
    d = torch.from_numpy(np.ascontiguousarray(box_deltas, dtype=np.float32)).cuda()
    return ops.decode_boxes(b, d, int(im_shape[0]), int(im_shape[1]), float(cfg.EPS)).cpu().numpy()


def _bbox_pred(boxes, box_deltas):
    """Unclipped decode (lib/detect/test.py:106-139)."""
    if boxes.shape[0] == 0:
        return np.zeros((0, box_deltas.shape[1]))
    b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float64)).cuda()
    d = torch.from_numpy(np.ascontiguousarray(box_deltas, dtype=np.float32)).cuda()
    out = ops.decode_boxes(b, d, 0, 0, float(cfg.EPS)).cpu().numpy()
    return out


def _clip_boxes(boxes, im_shape):
    """lib/detect/test.py:141-151 (in place, host: four vector min/max)."""
    boxes[:, 0::4] = np.maximum(boxes[:, 0::4], 0)
    boxes[:, 1::4] = np.maximum(boxes[:, 1::4], 0)
    boxes[:, 2::4] = np.minimum(boxes[:, 2::4], im_shape[1] - 1)
    boxes[:, 3::4] = np.minimum(boxes[:, 3::4], im_shape[0] - 1)
    return boxes


def divide_region(regions):
    """lib/detect/test.py:153-161 -> utils.cython_div.divide_region(regions, MIN_SIDE)."""
    return div.divide_region(np.ascontiguousarray(regions, dtype=np.float64), float(cfg.SEAR.MIN_SIDE))


def _unwrap_adj_pred(boxes, scores):
    """[R, 44] -> [11R, 4] roi-major, drop boxes whose shorter side (+1) is below MIN_SIDE (test.py:171-187)."""
    scores = scores.ravel()
    b = np.stack((boxes[:, 0::4].ravel(), boxes[:, 1::4].ravel(), boxes[:, 2::4].ravel(), boxes[:, 3::4].ravel()), axis=1)
    sides = np.minimum(b[:, 3] - b[:, 1] + 1, b[:, 2] - b[:, 0] + 1)
    keep = np.where(sides >= cfg.SEAR.MIN_SIDE)[0]
    return b[keep, :], scores[keep]


def _append_boxes(boxes):
    """SEAR.APPEND_BOXES (lib/detect/test.py:320-344): every proposal spawns the boxes of SEAR.APPEND_TEMP (itself, four
    boxes grown by a quarter on one side, one grown / one shrunk by an eighth all round), duplicates sifted on a grid of
    1 / DEDUP_BOXES pixels (`div._sift_dup`, CUDA).  Off by default (config.py:176)."""
    temp = cfg.SEAR.APPEND_TEMP                                   # [1, 4, n_templates]
    n, ns = boxes.shape[0], temp.shape[2]
    if n == 0:
        return np.zeros((0, 4))
    w, h = boxes[:, [2]] - boxes[:, [0]], boxes[:, [3]] - boxes[:, [1]]
    Lm = np.hstack((w, h, w, h))[:, :, np.newaxis]
    delta = np.hstack((boxes[:, [0]], boxes[:, [1]], boxes[:, [0]], boxes[:, [1]]))[:, :, np.newaxis]
    subs = np.transpose(Lm * temp + delta, [2, 0, 1]).reshape((n * ns, 4))
    return div._sift_dup(np.ascontiguousarray(subs, dtype=np.float64), 1 / cfg.DEDUP_BOXES)


def _dedup(rois_blob):
    v = np.array([1, 1e3, 1e6, 1e9, 1e12])
    hashes = np.round(rois_blob * cfg.DEDUP_BOXES).dot(v)
    _, index, inv = np.unique(hashes, return_index=True, return_inverse=True)
    return index, inv


def _net_forward(net, blobs, conv, conv_names):
    """Shared by _az_forward/_frcnn_forward: 'full' net on the first call, 'fc' net on the cached map."""
    if conv is None or 'fc' not in net.keys():
        net['full'].blobs['data'].reshape(*(blobs['data'].shape))
        net['full'].blobs['rois'].reshape(*(blobs['rois'].shape))
        out = net['full'].forward(data=blobs['data'].astype(np.float32, copy=False),
                                  rois=blobs['rois'].astype(np.float32, copy=False), blobs=conv_names)
        conv = {name: out[name] for name in conv_names}
        if 'fc' in net.keys() and isinstance(net['fc'], Net) and isinstance(net['full'], Net):
            net['fc'].adopt_conv(net['full'])
    else:
        for name in conv_names:
            net['fc'].blobs[name].reshape(*(conv[name].shape))
        net['fc'].blobs['rois'].reshape(*(blobs['rois'].shape))
        out = net['fc'].forward(rois=blobs['rois'].astype(np.float32, copy=False), **{n: conv[n] for n in conv_names})
    return out, conv


def _az_forward(net, im, all_boxes, conv=None):
    """Host-loop route of lib/detect/test.py:189-257 (one forward per BATCH_SIZE chunk)."""
    bs = cfg.SEAR.BATCH_SIZE
    z_all, a_all, c_all = np.zeros((0,)), np.zeros((0, 4)), np.zeros((0,))
    for start in range(0, all_boxes.shape[0], bs):
        boxes = all_boxes[start:start + bs, 0:4]
        blobs, _ = _get_blobs(im, boxes)
        inv = None
        if cfg.DEDUP_BOXES > 0:
            index, inv = _dedup(blobs['rois'])
            blobs['rois'] = blobs['rois'][index, :]
            boxes = boxes[index, :]
        # the 'full' net is asked for (and `conv` caches) SEAR.FRCNN_CONV, like the reference (test.py:222-226), so that
        # the shared maps can be handed to a skip-layer detector; the 'fc' net is fed SEAR.AZ_CONV (:231-236)
        full_pass = conv is None or 'fc' not in net.keys()
        out, conv = _net_forward(net, blobs, conv, cfg.SEAR.FRCNN_CONV if full_pass else cfg.SEAR.AZ_CONV)
        z = out['zoom_prob']
        scores = out['adj_prob']
        pred = _bbox_pred_clip(boxes, out['adj_bbox'], im.shape)
        if inv is not None:
            scores, pred, z = scores[inv, :], pred[inv, :], z[inv]
        a, c = _unwrap_adj_pred(pred, scores)
        z_all = np.hstack((z_all, z.ravel()))
        a_all = np.vstack((a_all, a))
        c_all = np.hstack((c_all, c))
    return z_all, a_all, c_all, conv


def _frcnn_forward(net, im, all_boxes, num_classes, conv=None):
    """lib/detect/test.py:259-318."""
    bs = cfg.SEAR.BATCH_SIZE
    all_pred, all_scores = np.zeros((0, 4 * num_classes)), np.zeros((0, num_classes))
    for start in range(0, all_boxes.shape[0], bs):
        boxes = all_boxes[start:start + bs, 0:4]
        blobs, _ = _get_blobs(im, boxes)
        inv = None
        if cfg.DEDUP_BOXES > 0:
            index, inv = _dedup(blobs['rois'])
            blobs['rois'] = blobs['rois'][index, :]
            boxes = boxes[index, :]
        out, conv = _net_forward(net, blobs, conv, cfg.SEAR.FRCNN_CONV)
        scores = out['cls_prob']
        pred = _bbox_pred_clip(boxes, out['bbox_pred'], im.shape)
        if inv is not None:
            scores, pred = scores[inv, :], pred[inv, :]
        all_scores = np.vstack((all_scores, scores))
        all_pred = np.vstack((all_pred, pred))
    return all_scores, all_pred, conv


# --------------------------------------------------------------------------- the adaptive search
def _engine_for(az_net: Net, im_shape, num_proposals, n_img=1):
    return batched.search_engine(az_net, im_shape, n_img, num_proposals)


def _fast_route(net):
    full = net.get('full') if hasattr(net, 'get') else None
    return isinstance(full, Net) and full.kind == "az" and full.backbone is not None and len(cfg.TEST.SCALES) == 1


def _device_maps(full: Net, im, names):
    """One image -> the bf16 NHWC maps `names` of one backbone pass, network input built on the device
    (azn_image_blob: the mean-subtract + cv2.resize of _get_image_blob, test.py:27-59, without the host)."""
    pix = torch.from_numpy(np.ascontiguousarray(im, dtype=np.uint8)[None]).to(full.dev)
    scale = im_scale_for(im.shape[0], im.shape[1], tuple(cfg.TEST.SCALES), cfg.TEST.MAX_SIZE)
    bb = full.backbone
    return bb.run_padded(ops.image_blob(pix, scale, bb.pixel_means, bb.cpad_in), taps=tuple(names))


def im_propose(net, im, return_conv=False, num_proposals=None):
    """Generate object proposals with AZ-Net (lib/detect/test.py:346-414).
    net: {'full': Net[, 'fc': Net]}; im: HxWx3 uint8 BGR.  Returns Y [n,4] float64 (and conv dict)."""
    if _fast_route(net):
        full = net['full']
        eng = _engine_for(full, im.shape, num_proposals)
        names = tuple(cfg.SEAR.FRCNN_CONV)
        taps = None
        if return_conv and names != ('conv5_3',):
            # skip-layer detector (experiments/cfgs/voc_skip.yml:20): hand out conv3_3 / conv4_3 / conv5_3 of the same pass
            taps = _device_maps(full, im, tuple(dict.fromkeys(names + ('conv5_3',))))
            nhwc = taps['conv5_3']
        else:
            nhwc = _device_maps(full, im, ('conv5_3',))['conv5_3']
        eng.propose(nhwc)
        boxes, _, n_eval, depth = eng.results()
        Y, num_eval, k = boxes[0], int(n_eval[0]), int(depth[0])
        conv = None
        if return_conv and taps is not None:
            conv = {}
            for name in names:
                dev = taps[name].permute(0, 3, 1, 2).float().contiguous()
                host = dev.cpu().numpy()
                full._maps[name] = (full._map_key(host), dev, taps[name], host)
                conv[name] = host
        elif return_conv:
            conv_dev = nhwc.permute(0, 3, 1, 2).float().contiguous()
            host = conv_dev.cpu().numpy()
            full._last_conv = (host, conv_dev, nhwc)
            conv = {name: host for name in cfg.SEAR.FRCNN_CONV}
    else:
        B = np.array([[0, 0, im.shape[1] - 1.0, im.shape[0] - 1.0]])
        Y, a_scores = np.zeros((0, 4)), np.zeros((0,))
        num_eval, conv, k = 0, None, 0
        for k in range(1, search_depth(im.shape[0], im.shape[1], cfg.SEAR.MIN_SIDE)):
            zoom, boxes, c, conv = _az_forward(net, im, B, conv)
            num_eval += B.shape[0]
            Y = np.vstack((Y, boxes))
            a_scores = np.hstack((a_scores, c))
            if k == 1:
                zoom[0] = 1.0                         # the root region is always divided
            Z = B[np.where(zoom >= cfg.SEAR.Tz)[0], :]
            if Z.shape[0] == 0:
                break
            B = divide_region(Z)
        if (not cfg.SEAR.FIXED_PROPOSAL_NUM) and (num_proposals is None):
            Y = Y[np.where(a_scores >= cfg.SEAR.Tc)[0], :]
        else:
            n = cfg.SEAR.NUM_PROPOSALS if num_proposals is None else num_proposals
            Y = Y[np.argsort(-a_scores, kind='stable')[:min(n, Y.shape[0])], :]
    if cfg.SEAR.APPEND_BOXES:                                     # test.py:403-406
        Y = _clip_boxes(_append_boxes(Y), im.shape)
    print('{0} proposals, evaluate {1} regions, reaches depth {2}.'.format(Y.shape[0], num_eval, k))
    return (Y, conv) if return_conv else Y


def im_propose_batch(net, conv_maps, im_shape, num_proposals=None, engine=None):
    """Batched search over same-sized images whose conv5_3 maps are already computed.
    conv_maps: f32 NCHW host/device tensor or ndarray [n, C, h, w].  Returns (list of Y, list of scores)."""
    az = net['fc'] if 'fc' in net.keys() else net['full']
    maps = conv_maps if torch.is_tensor(conv_maps) else torch.from_numpy(np.ascontiguousarray(conv_maps, dtype=np.float32))
    n = maps.shape[0]
    if engine is None:
        fixed = bool(cfg.SEAR.FIXED_PROPOSAL_NUM) or num_proposals is not None
        engine = SearchEngine(az.head, n, im_shape[0], im_shape[1], scales=tuple(cfg.TEST.SCALES), max_size=cfg.TEST.MAX_SIZE,
                              min_side=cfg.SEAR.MIN_SIDE, tz=float(cfg.SEAR.Tz), tc=float(cfg.SEAR.Tc), fixed_num=fixed,
                              num_proposals=num_proposals if num_proposals is not None else cfg.SEAR.NUM_PROPOSALS,
                              batch_size=cfg.SEAR.BATCH_SIZE, dedup=float(cfg.DEDUP_BOXES), eps=float(cfg.EPS))
    engine.propose(ops.nchw_to_nhwc_bf16(maps.to(engine.dev, non_blocking=True)))
    boxes, scores, _, _ = engine.results()
    return boxes, scores


def im_detect(net, im, boxes, num_classes):
    """Fast R-CNN scores [R,C] and class-specific boxes [R,4C] for the proposals (test.py:416-430)."""
    scores, pred_boxes, _ = _frcnn_forward(net, im, boxes, num_classes)
    return scores, pred_boxes


def im_detect_shared(az_net, frcnn_net, im, num_classes):
    """AZ-Net proposals + Fast R-CNN on the shared conv map (test.py:432-445)."""
    boxes, conv = im_propose(az_net, im, return_conv=True)
    if isinstance(frcnn_net.get('fc'), Net) and isinstance(az_net.get('full'), Net):
        frcnn_net['fc'].adopt_conv(az_net['full'])
    scores, pred_boxes, _ = _frcnn_forward(frcnn_net, im, boxes, num_classes, conv)
    return scores, pred_boxes


def apply_nms(all_boxes, thresh):
    """NMS over all_boxes[cls][img] (test.py:467-484): every (class, image) problem of the whole set goes
    into ONE batched kernel launch (segments above AZN_NMS_SEG_MAX rows use the large-N path)."""
    from .. import _lib as L
    num_classes, num_images = len(all_boxes), len(all_boxes[0])
    out = [[[] for _ in range(num_images)] for _ in range(num_classes)]
    small, big = [], []
    for c in range(num_classes):
        for i in range(num_images):
            d = all_boxes[c][i]
            if isinstance(d, list) and d == []:
                continue
            if not isinstance(d, np.ndarray) or d.dtype != np.float32 or d.ndim != 2:
                raise ValueError("Buffer dtype mismatch, expected 'float32_t'")
            if d.shape[0] == 0:
                continue
            (small if d.shape[0] <= L.NMS_SEG_MAX else big).append((c, i, d))
    if small:
        sizes = np.array([d.shape[0] for _, _, d in small], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        dets = torch.from_numpy(np.ascontiguousarray(np.vstack([d[:, :5] for _, _, d in small]))).cuda()
        keep, cnt = ops.nms_batched(dets, torch.from_numpy(off).cuda(), float(thresh))
        keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
        for s, (c, i, d) in enumerate(small):
            k = keep[off[s]:off[s] + cnt[s]]
            if len(k):
                out[c][i] = d[k, :].copy()
    for c, i, d in big:
        k = nms(d, float(thresh))
        if len(k):
            out[c][i] = d[k, :].copy()
    return out


# --------------------------------------------------------------------------- device-resident detection
def _fast_detect_route(net, key='full'):
    n = net.get(key) if hasattr(net, 'get') else None
    return (isinstance(n, Net) and n.kind in ("frcnn", "frcnn_skip") and len(cfg.TEST.SCALES) == 1
            and (key == 'fc' or n.backbone is not None))


def _boxes_to_device(prop_list, cap, dev):
    """Per-image proposal arrays (f64 [R_i, >=4]) -> zero-padded device tensor [n, cap, 4] + int32 counts."""
    n = len(prop_list)
    host = np.zeros((n, cap, 4))
    counts = np.zeros((n,), np.int32)
    for k, b in enumerate(prop_list):
        if b is None or len(b) == 0:
            continue
        counts[k] = b.shape[0]
        host[k, :b.shape[0]] = b[:, :4]
    return torch.from_numpy(host).to(dev), torch.from_numpy(counts).to(dev)


def _proposal_cap(max_rows):
    """Row capacity of the detector's proposal buffer: the configured proposal count (SEAR.NUM_PROPOSALS exists only
    after cfg_set_mode -- tools/test_det_net.py never calls it -- so fall back to TEST.NUM_PROPOSALS), or more when a
    proposals.pkl holds longer lists."""
    want = int(getattr(cfg.SEAR, 'NUM_PROPOSALS', cfg.TEST.NUM_PROPOSALS))
    return max(want, (int(max_rows) + 127) // 128 * 128)


def _finish_on_device(dset: DetectionSet, skip, imdb, output_dir, shard=None):
    """thresholds + final filter + NMS for the whole set on the device; same files and calls as _finish_detections.
    shard = (rank, world, range of this rank's images): `dset` then holds this rank's images only; its finish() exchanges
    the scores so that every rank computes the set-wide thresholds (dist.gather_detection_scores), the per-image results
    are merged into the database's nesting on every rank and rank 0 writes the file and evaluates."""
    rank, world, mine = shard if shard is not None else (0, 1, range(len(imdb.image_index)))
    dset.finish(cfg.TEST.NMS)

    def whole(local):
        if world == 1:
            return local
        import torch.distributed as dist
        parts = [None] * world
        dist.all_gather_object(parts, (mine.start, [[local[j][k] for k in range(len(mine))] for j in range(imdb.num_classes)]))
        out = [[[] for _ in range(len(imdb.image_index))] for _ in range(imdb.num_classes)]
        for lo, part in parts:
            for j in range(imdb.num_classes):
                for k, d in enumerate(part[j]):
                    out[j][lo + k] = d
        return out

    all_boxes = whole(dset.to_host())
    for j in range(imdb.num_classes):
        for i in skip:
            all_boxes[j][i] = []
    if rank == 0:
        with open(os.path.join(output_dir, 'detections.pkl'), 'wb') as f:
            pickle.dump(all_boxes, f, pickle.HIGHEST_PROTOCOL)
    print('Applying NMS to all detections')
    nms_dets = whole(dset.to_host(nms=True))
    print('Evaluating detections')
    if rank == 0:
        imdb.evaluate_detections(nms_dets, output_dir)


class _BatchClock:
    """The reference's per-image Timer (utils/timer.py) for batched execution: average_time = wall time since the
    first batch was submitted / images finished."""

    def __init__(self):
        self.timer, self.t0 = Timer(), None

    def start(self):
        import time
        if self.t0 is None:
            self.t0 = time.time()

    def done(self, n_images):
        import time
        t = self.timer
        t.calls += n_images
        t.total_time = time.time() - self.t0
        t.average_time = t.total_time / max(t.calls, 1)
        return t.average_time


# --------------------------------------------------------------------------- dataset drivers
def _image_shard(num_images):
    """Under torch.distributed (one process per GPU, e.g. `torchrun tools/prop_az.py`) the image database is cut into
    contiguous blocks, one per rank (aznet_b200.dist.shard_images): images are independent (test.py:508-513).
    -> (rank, world, range of this rank's image indices)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        from ..dist import shard_images
        rank, world = dist.get_rank(), dist.get_world_size()
        lo, hi = shard_images(num_images, rank, world)
        return rank, world, range(lo, hi)
    return 0, 1, range(num_images)


def _merge_image_lists(prop_boxes, indices, world):
    """Every rank's per-image lists -> the whole database's list on every rank (one object all_gather at the end of the
    job: the lists are host arrays by now, 300 x 4 f64 per image)."""
    if world == 1:
        return
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, {i: prop_boxes[i] for i in indices})
    for part in parts:
        for i, b in part.items():
            prop_boxes[i] = b


def _test_proposals_batched(net, imdb, prop_boxes, stats, indices=None):
    """The fast route of test_proposals: read-ahead, same-shape batches, device image blob, batched backbone + search
    (detect/batched.py).  Prints the reference's per-image lines as each batch finishes."""
    full = net['full']
    indices = range(len(imdb.image_index)) if indices is None else indices
    num_images = len(indices)
    copy_stream = torch.cuda.Stream(device=full.dev)
    clock = _BatchClock()
    done = [0]

    def finish(item):
        db, fetch = item
        r = fetch.get()
        if int(r['status'][0]) != 0:
            raise RuntimeError("search capacity overflow on device (status %d)" % int(r['status'][0]))
        avg = clock.done(db.n)
        for k, i in enumerate(db.idx):
            c = int(r['count'][k])
            prop_boxes[i] = r['boxes'][k, :c].copy()
            if cfg.SEAR.APPEND_BOXES:                             # test.py:403-406, per image on the fetched list
                prop_boxes[i] = _clip_boxes(_append_boxes(prop_boxes[i]), db.shape)
                c = prop_boxes[i].shape[0]
            stats['num_eval'] += int(r['n_eval'][k])
            print('{0} proposals, evaluate {1} regions, reaches depth {2}.'.format(c, int(r['n_eval'][k]), int(r['depth'][k])))
            done[0] += 1
            print('im_prop: {:d}/{:d} {:.3f}s'.format(done[0], num_images, avg))

    import time
    pending = None
    t_mark = time.perf_counter()
    for batch in batched.ImageFeeder(imdb, indices):
        t0 = time.perf_counter()
        stats['host_feed_s'] = stats.get('host_feed_s', 0.0) + (t0 - t_mark)          # read-ahead + pinned staging
        clock.start()
        db = batched.DeviceBatch(full, batch, copy_stream)
        eng = _engine_for(full, db.shape + (3,), None, db.n_pad)
        eng.propose(db.maps)
        fetch = batched.Fetch(boxes=eng.out_boxes[:db.n], count=eng.out_count[:db.n], n_eval=eng.n_eval[:db.n],
                              depth=eng.depth[:db.n], status=eng.status)
        stats['h2d_bytes'] += db.h2d_bytes
        stats['d2h_bytes'] += fetch.bytes
        stats['batches'] += 1
        t1 = time.perf_counter()
        stats['host_submit_s'] = stats.get('host_submit_s', 0.0) + (t1 - t0)           # launches of one batch
        if pending is not None:
            finish(pending)
        pending = (db, fetch)
        t_mark = time.perf_counter()
        stats['host_finish_s'] = stats.get('host_finish_s', 0.0) + (t_mark - t1)       # waiting for the previous batch + unpack
    if pending is not None:
        t1 = time.perf_counter()
        finish(pending)
        stats['host_finish_s'] = stats.get('host_finish_s', 0.0) + (time.perf_counter() - t1)
    return clock.timer


def test_proposals(net, imdb):
    """Generate proposals on an image database and pickle them (test.py:486-539)."""
    import cv2
    num_images = len(imdb.image_index)
    prop_boxes = [[] for _ in range(num_images)]
    output_dir = get_output_dir(imdb, net['full'])
    os.makedirs(output_dir, exist_ok=True)
    num_boxes = 0.0
    stats = test_proposals.last_stats = {'route': 'host', 'num_eval': 0, 'h2d_bytes': 0, 'd2h_bytes': 0, 'batches': 0}
    # one process per GPU: this rank's block of the images; the lists are merged at the end and rank 0 writes the file
    rank, world, mine = _image_shard(num_images)
    stats['rank'], stats['world'], stats['images'] = rank, world, len(mine)
    if _fast_route(net):
        stats['route'] = 'batched'
        t = _test_proposals_batched(net, imdb, prop_boxes, stats, mine)
    else:
        t = Timer()
        for k, i in enumerate(mine):
            im = cv2.imread(imdb.image_path_at(i))
            t.tic()
            prop_boxes[i] = im_propose(net, im)
            t.toc()
            print('im_prop: {:d}/{:d} {:.3f}s'.format(k + 1, len(mine), t.average_time))
    _merge_image_lists(prop_boxes, mine, world)
    recall = 0
    prop = {'boxes': prop_boxes, 'time': t.average_time, 'recall': recall}
    if rank == 0:
        with open(os.path.join(output_dir, 'proposals.pkl'), 'wb') as f:
            pickle.dump(prop, f, pickle.HIGHEST_PROTOCOL)
    print('The recall is {:.3f}'.format(recall))
    print('On average, {0} boxes per image are generated'.format(num_boxes / num_images))
    print('The average proposal generation time is {:.3f}s'.format(t.average_time))


def _select_detections(scores, boxes, thresh, top_scores, max_per_image, max_per_set, all_boxes, i, num_classes):
    """Per-class selection for one image (test.py:608-636): score > thresh[j], top-100, heap cap."""
    for j in range(1, num_classes):
        inds = np.where(scores[:, j] > thresh[j])[0]
        cls_scores = scores[inds, j]
        cls_boxes = boxes[inds, j * 4:(j + 1) * 4]
        top = np.argsort(-cls_scores, kind='stable')[:max_per_image]
        cls_scores, cls_boxes = cls_scores[top], cls_boxes[top, :]
        for val in cls_scores:
            heapq.heappush(top_scores[j], val)
        if len(top_scores[j]) > max_per_set:
            while len(top_scores[j]) > max_per_set:
                heapq.heappop(top_scores[j])
            thresh[j] = top_scores[j][0]
        all_boxes[j][i] = np.hstack((cls_boxes, cls_scores[:, np.newaxis])).astype(np.float32, copy=False)


def _finish_detections(all_boxes, thresh, skip, imdb, output_dir):
    for j in range(1, imdb.num_classes):
        for i in range(len(imdb.image_index)):
            if i in skip:
                continue
            inds = np.where(all_boxes[j][i][:, -1] > thresh[j])[0]
            all_boxes[j][i] = all_boxes[j][i][inds, :]
    with open(os.path.join(output_dir, 'detections.pkl'), 'wb') as f:
        pickle.dump(all_boxes, f, pickle.HIGHEST_PROTOCOL)
    print('Applying NMS to all detections')
    nms_dets = apply_nms(all_boxes, cfg.TEST.NMS)
    print('Evaluating detections')
    imdb.evaluate_detections(nms_dets, output_dir)


def _detect_batch(fr_net, db, boxes_d, count_d, dset, copy_back=True, base=0):
    """One batch's Fast R-CNN step (DetectEngine over all images of the batch) into the set-wide buffers at the
    batch's image indices (minus `base`: the first image of this rank's shard)."""
    eng = batched.detect_engine(fr_net, db.shape, db.n_pad, boxes_d.shape[1])
    dets, tops, cnt = eng.detect(db.maps, boxes_d, count_d)
    idx = torch.tensor([i - base for i in db.idx], dtype=torch.long, device=fr_net.dev)
    dset.dets.index_copy_(0, idx, dets[:db.n])
    dset.top_scores.index_copy_(0, idx, tops[:db.n])
    dset.det_count.index_copy_(0, idx, cnt[:db.n])
    return eng


def _test_net_batched(net, prop_boxes, imdb, dset, todo, stats, base=0, num_images=None):
    """Fast route of test_net: batches of same-shape images, backbone + DetectEngine per batch, nothing but the
    uint8 pixels and the proposal lists crosses PCIe before the set-wide finish."""
    full = net['full']
    num_images = len(imdb.image_index) if num_images is None else num_images
    copy_stream = torch.cuda.Stream(device=full.dev)
    clock = _BatchClock()
    done, num_boxes = 0, 0.0
    cap = _proposal_cap(max([prop_boxes[i].shape[0] for i in todo] + [1]))
    pending = None
    for batch in batched.ImageFeeder(imdb, todo):
        clock.start()
        db = batched.DeviceBatch(full, batch, copy_stream, taps=full.conv_names)
        plist = [prop_boxes[i] for i in db.idx] + [None] * (db.n_pad - db.n)
        boxes_d, count_d = _boxes_to_device(plist, cap, full.dev)
        _detect_batch(full, db, boxes_d, count_d, dset, base=base)
        ev = torch.cuda.Event()
        ev.record()
        stats['h2d_bytes'] += db.h2d_bytes + boxes_d.numel() * 8
        stats['batches'] += 1
        if pending is not None:
            pdb, pev = pending
            pev.synchronize()
            avg = clock.done(pdb.n)
            for i in pdb.idx:
                done += 1
                num_boxes += prop_boxes[i].shape[0]
                print('im_detect: {:d}/{:d} {:.3f}s {:.3f}s'.format(done, num_images, avg, 0.0))
        pending = (db, ev)
    if pending is not None:
        pdb, pev = pending
        pev.synchronize()
        avg = clock.done(pdb.n)
        for i in pdb.idx:
            done += 1
            num_boxes += prop_boxes[i].shape[0]
            print('im_detect: {:d}/{:d} {:.3f}s {:.3f}s'.format(done, num_images, avg, 0.0))
    return clock.timer, num_boxes


def test_net(net, prop_file, imdb):
    """Fast R-CNN over pre-computed proposals (test.py:541-668)."""
    import cv2
    with open(prop_file, 'rb') as f:
        prop = pickle.load(f)
    prop_boxes = prop['boxes']
    num_images = len(imdb.image_index)
    max_per_set = 800 // (imdb.num_classes - 1) * num_images       # Python-2 integer division
    max_per_image = 100
    thresh = -np.inf * np.ones(imdb.num_classes)
    top_scores = [[] for _ in range(imdb.num_classes)]
    all_boxes = [[[] for _ in range(num_images)] for _ in range(imdb.num_classes)]
    num_boxes = 0.0
    output_dir = get_output_dir(imdb, net['full'])
    os.makedirs(output_dir, exist_ok=True)
    _t = {'im_detect': Timer(), 'misc': Timer()}
    skip = set(i for i in range(num_images) if prop_boxes[i].shape[0] == 0)
    stats = test_net.last_stats = {'route': 'host', 'h2d_bytes': 0, 'batches': 0}
    if _fast_detect_route(net):
        stats['route'] = 'batched'
        # one process per GPU: this rank's block of the images; the thresholds are set-wide (one exchange of the scores)
        shard = _image_shard(num_images)
        rank, world, mine = shard
        stats['rank'], stats['world'], stats['images'] = rank, world, len(mine)
        dset = DetectionSet(len(mine), imdb.num_classes, max_per_image, net['full'].dev, total_images=num_images)
        _t['im_detect'], num_boxes = _test_net_batched(net, prop_boxes, imdb, dset, [i for i in mine if i not in skip], stats,
                                                       base=mine.start, num_images=len(mine))
        _finish_on_device(dset, skip, imdb, output_dir, shard)
    else:
        for i in range(num_images):
            if i in skip:
                continue
            im = cv2.imread(imdb.image_path_at(i))
            _t['im_detect'].tic()
            scores, boxes = im_detect(net, im, prop_boxes[i], imdb.num_classes)
            num_boxes += scores.shape[0]
            _t['im_detect'].toc()
            _t['misc'].tic()
            _select_detections(scores, boxes, thresh, top_scores, max_per_image, max_per_set, all_boxes, i, imdb.num_classes)
            _t['misc'].toc()
            print('im_detect: {:d}/{:d} {:.3f}s {:.3f}s'.format(i + 1, num_images, _t['im_detect'].average_time,
                                                                _t['misc'].average_time))
        _finish_detections(all_boxes, thresh, skip, imdb, output_dir)
    print('The average time is proposal {:.3f}s, detection {:.3f}s'.format(prop['time'], _t['im_detect'].average_time))
    print('On average, {0} boxes per image are generated'.format(num_boxes / num_images))


def _test_net_shared_batched(sc_net, frcnn_net, imdb, dset, stats, indices=None):
    """Fast route of test_net_shared: one backbone pass per batch serves the search and the detector; the proposals
    never leave the device (the search engine's output buffers are the detector's input)."""
    full, fr = sc_net['full'], frcnn_net['fc']
    indices = range(len(imdb.image_index)) if indices is None else indices
    num_images = len(indices)
    copy_stream = torch.cuda.Stream(device=full.dev)
    clock = _BatchClock()
    state = {'done': 0, 'boxes': 0.0}
    taps = tuple(dict.fromkeys(tuple(fr.conv_names) + ('conv5_3',)))

    def finish(item):
        db, fetch = item
        r = fetch.get()
        if int(r['status'][0]) != 0:
            raise RuntimeError("search capacity overflow on device (status %d)" % int(r['status'][0]))
        avg = clock.done(db.n)
        for k in range(db.n):
            print('{0} proposals, evaluate {1} regions, reaches depth {2}.'.format(int(r['count'][k]), int(r['n_eval'][k]),
                                                                                   int(r['depth'][k])))
            state['boxes'] += int(r['count'][k])
            state['done'] += 1
            print('im_detect: {:d}/{:d} {:.3f}s {:.3f}s'.format(state['done'], num_images, avg, 0.0))

    pending = None
    for batch in batched.ImageFeeder(imdb, indices):
        clock.start()
        db = batched.DeviceBatch(full, batch, copy_stream, taps=taps)
        maps = db.maps
        conv5 = maps['conv5_3'] if isinstance(maps, dict) else maps
        eng = _engine_for(full, db.shape + (3,), None, db.n_pad)
        eng.propose(conv5)
        if db.n_pad > db.n:
            eng.out_count[db.n:].zero_()                      # padding images propose nothing
        if fr.kind != "frcnn_skip":
            db.maps = conv5
        _detect_batch(fr, db, eng.out_boxes, eng.out_count, dset, base=indices[0] if len(indices) else 0)
        fetch = batched.Fetch(count=eng.out_count[:db.n], n_eval=eng.n_eval[:db.n], depth=eng.depth[:db.n], status=eng.status)
        stats['h2d_bytes'] += db.h2d_bytes
        stats['batches'] += 1
        if pending is not None:
            finish(pending)
        pending = (db, fetch)
    if pending is not None:
        finish(pending)
    return clock.timer, state['boxes']


def test_net_shared(sc_net, frcnn_net, imdb):
    """Shared-conv detection (test.py:670-780): proposals and detection in one pass per image."""
    import cv2
    num_images = len(imdb.image_index)
    max_per_set = 800 // (imdb.num_classes - 1) * num_images
    max_per_image = 100
    thresh = -np.inf * np.ones(imdb.num_classes)
    top_scores = [[] for _ in range(imdb.num_classes)]
    all_boxes = [[[] for _ in range(num_images)] for _ in range(imdb.num_classes)]
    num_boxes = 0.0
    output_dir = get_output_dir(imdb, sc_net['full'])
    os.makedirs(output_dir, exist_ok=True)
    _t = {'im_detect': Timer(), 'misc': Timer()}
    stats = test_net_shared.last_stats = {'route': 'host', 'h2d_bytes': 0, 'batches': 0}
    if _fast_route(sc_net) and _fast_detect_route(frcnn_net, 'fc') and not cfg.SEAR.APPEND_BOXES:
        stats['route'] = 'batched'
        shard = _image_shard(num_images)
        rank, world, mine = shard
        stats['rank'], stats['world'], stats['images'] = rank, world, len(mine)
        dset = DetectionSet(len(mine), imdb.num_classes, max_per_image, sc_net['full'].dev, total_images=num_images)
        _t['im_detect'], num_boxes = _test_net_shared_batched(sc_net, frcnn_net, imdb, dset, stats, mine)
        _finish_on_device(dset, set(), imdb, output_dir, shard)
    else:
        for i in range(num_images):
            im = cv2.imread(imdb.image_path_at(i))
            _t['im_detect'].tic()
            scores, boxes = im_detect_shared(sc_net, frcnn_net, im, imdb.num_classes)
            num_boxes += scores.shape[0]
            _t['im_detect'].toc()
            _t['misc'].tic()
            _select_detections(scores, boxes, thresh, top_scores, max_per_image, max_per_set, all_boxes, i, imdb.num_classes)
            _t['misc'].toc()
            print('im_detect: {:d}/{:d} {:.3f}s {:.3f}s'.format(i + 1, num_images, _t['im_detect'].average_time,
                                                                _t['misc'].average_time))
        _finish_detections(all_boxes, thresh, set(), imdb, output_dir)
    print('The average detection time is {:.3f}s'.format(_t['im_detect'].average_time))
    print('On average, {0} boxes per image are proposed'.format(num_boxes / num_images))
