"""Drop-in for lib/detect/tune.py: the diagnostic search that records every anchor region with its zoom
score, and the calibration of the zoom threshold -- what tools/set_thresh.py (:94) and tools/diagnose_prop.py
(:102) call.  Produces the `thresh.pkl` wire format that cfg_load_thresh / tools/prop_az.py consume
(lib/detect/config.py:282-287), so the proposal pipeline is self-contained.

    im_propose(net, im)        -> (hstack(Y, scores) [n,5], Bhis [m,5])       tune.py:256-316
    tune_thresh(net, imdb)     -> writes <output_dir>/thresh.pkl              tune.py:318-366
    test_proposals(net, imdb)  -> writes <output_dir>/AZ_results.mat          tune.py:368-419

Differences of this im_propose from detect.test.im_propose (all the reference's): K levels instead of K-1
(`for k in xrange(K)`), the root is not forced -- the first level compares its zoom score with Tz = 0 -- and the
proposals come back with their scores.  With aznet_b200.net.Net the whole loop runs device-resident in
SearchEngine(tune=True) (history written by azn_search_level); the threshold over the image set is
azn_tune_threshold (the reference's min-heap is order-independent).  No CPU fallback.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from .config import cfg, get_output_dir
from . import test as _t
from .. import ops
from ..engine import SearchEngine, search_depth
from ..utils.timer import Timer

_ENGINES = {}


def _engine_for(az_net, im_shape):
    key = (id(az_net.head), int(im_shape[0]), int(im_shape[1]), tuple(cfg.TEST.SCALES), cfg.TEST.MAX_SIZE, cfg.SEAR.MIN_SIDE,
           float(cfg.SEAR.Tz), cfg.SEAR.NUM_PROPOSALS, cfg.SEAR.BATCH_SIZE, float(cfg.DEDUP_BOXES), float(cfg.EPS))
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) > 16:
            _ENGINES.clear()
        eng = SearchEngine(az_net.head, 1, im_shape[0], im_shape[1], scales=tuple(cfg.TEST.SCALES), max_size=cfg.TEST.MAX_SIZE,
                           min_side=cfg.SEAR.MIN_SIDE, tz=float(cfg.SEAR.Tz), fixed_num=True,
                           num_proposals=cfg.SEAR.NUM_PROPOSALS, batch_size=cfg.SEAR.BATCH_SIZE,
                           dedup=float(cfg.DEDUP_BOXES), eps=float(cfg.EPS), spatial_scale=az_net.spatial_scale, tune=True)
        _ENGINES[key] = eng
    return eng


def _propose(net, im, keep_on_device=False):
    """-> (Y5, Bhis, engine or None).  Y5 f64 [n,5], Bhis f64 [m,5] (region, zoom score)."""
    if _t._fast_route(net):
        full = net['full']
        eng = _engine_for(full, im.shape)
        data, _ = _t._get_image_blob(im)
        _, nhwc = full.conv_from_data(data)
        eng.propose(nhwc)
        boxes, scores, n_eval, depth = eng.results()
        Y5 = np.hstack((boxes[0], scores[0][:, np.newaxis].astype(np.float64)))
        num_eval, k = int(n_eval[0]), int(depth[0]) - 1
        Bhis = None if keep_on_device else eng.history()[0]
    else:
        eng = None
        B = np.array([[0, 0, im.shape[1] - 1.0, im.shape[0] - 1.0]])
        Bhis = np.zeros((0, 5))
        Y, a_scores = np.zeros((0, 4)), np.zeros((0,))
        num_eval, conv, k, Tz = 0, None, 0, 0
        for k in range(search_depth(im.shape[0], im.shape[1], cfg.SEAR.MIN_SIDE)):
            zoom, boxes, c, conv = _t._az_forward(net, im, B, conv)
            num_eval += B.shape[0]
            Y = np.vstack((Y, boxes))
            a_scores = np.hstack((a_scores, c))
            Z = B[np.where(zoom >= Tz)[0], :]
            Bhis = np.vstack((Bhis, np.hstack((B, zoom[:, np.newaxis]))))
            if Z.shape[0] == 0:
                break
            B = _t.divide_region(Z)
            Tz = cfg.SEAR.Tz
        ind = np.argsort(-a_scores, kind='stable')[:min(cfg.SEAR.NUM_PROPOSALS, Y.shape[0])]
        Y5 = np.hstack((Y[ind, :], a_scores[ind, np.newaxis]))
    print('{0} proposals, evaluate {1} regions, reaches depth {2}.'.format(Y5.shape[0], num_eval, k))
    return Y5, Bhis, eng


def im_propose(net, im):
    """Proposals with scores + the anchor-region history (lib/detect/tune.py:256-316)."""
    Y5, Bhis, _ = _propose(net, im)
    return Y5, Bhis


def tune_thresh(net, imdb):
    """Find the zoom threshold for which the average number of anchors per image is about
    cfg.TRAIN.ANCHORS_PER_IMG, and pickle it (lib/detect/tune.py:318-366)."""
    import cv2
    num_images = len(imdb.image_index)
    max_per_set = num_images * cfg.TRAIN.ANCHORS_PER_IMG
    output_dir = get_output_dir(imdb, net['full'])
    if not os.path.exists(output_dir):
        os.makedirs(output_dir)
    t = Timer()
    zooms = []                                     # per image: f32 zoom scores of its anchors, on the device
    for i in range(num_images):
        im = cv2.imread(imdb.image_path_at(i))
        t.tic()
        _, Bhis, eng = _propose(net, im, keep_on_device=True)
        if eng is not None:
            zooms.append(eng.hist_zoom[0, :int(eng.n_history[0].item())].clone())
        else:
            zooms.append(torch.from_numpy(np.ascontiguousarray(Bhis[:, -1], dtype=np.float32)).cuda())
        t.toc()
        print('im_tune: {:d}/{:d} {:.3f}s'.format(i + 1, num_images, t.average_time))
    cap = max(max(int(z.numel()) for z in zooms), 1)
    table = torch.zeros((num_images, cap), dtype=torch.float32, device=zooms[0].device)
    for i, z in enumerate(zooms):
        table[i, :z.numel()] = z
    counts = torch.tensor([int(z.numel()) for z in zooms], dtype=torch.int32, device=table.device)
    th = float(ops.tune_threshold(table, counts, max_per_set).item())
    thresh = np.float64(th) if np.isfinite(th) else -np.inf
    print('the threshold is set to {0}'.format(thresh))
    with open(os.path.join(output_dir, 'thresh.pkl'), 'wb') as f:
        pickle.dump(thresh, f, pickle.HIGHEST_PROTOCOL)
    return thresh


def test_proposals(net, imdb):
    """Record proposals, anchor regions and ground truth of every image for fine-grained analysis
    (lib/detect/tune.py:368-419): <output_dir>/AZ_results.mat."""
    import cv2
    import scipy.io as sio
    num_images = len(imdb.image_index)
    prop_boxes = np.zeros((num_images,), dtype=object)
    anchor_boxes = np.zeros((num_images,), dtype=object)
    gt_boxes = np.zeros((num_images,), dtype=object)
    fn = np.zeros((num_images,), dtype=object)
    im_shapes = np.zeros((num_images,), dtype=object)
    output_dir = get_output_dir(imdb, net['full'])
    if not os.path.exists(output_dir):
        os.makedirs(output_dir)
    t = Timer()
    gt_roidb = imdb.gt_roidb()
    for i in range(num_images):
        im = cv2.imread(imdb.image_path_at(i))
        im_shapes[i] = im.shape
        t.tic()
        prop_boxes[i], anchor_boxes[i] = im_propose(net, im)
        t.toc()
        gt_boxes[i] = gt_roidb[i]['boxes']
        fn[i] = os.path.basename(imdb.image_path_at(i))
        print('im_prop: {:d}/{:d} {:.3f}s'.format(i + 1, num_images, t.average_time))
    sio.savemat(os.path.join(output_dir, 'AZ_results.mat'),
                dict(prop_boxes=prop_boxes, anchor_boxes=anchor_boxes, gt_boxes=gt_boxes, fn=fn, Tz=cfg.SEAR.Tz,
                     num_proposals=cfg.SEAR.NUM_PROPOSALS, im_shapes=im_shapes))
