"""AZ-detect config system: the host-side mirror of lib/detect/config.py (same `cfg` tree, key names,
default values and setter functions), so that tools/prop_az.py / tools/test_det_net.py style drivers work
unchanged:  cfg, cfg_from_file, cfg_set_mode, cfg_load_thresh, cfg_set_path, get_output_dir.

`easydict` is not installed in this image; EasyDict below is a minimal stand-in with attribute access.
The training-only keys are kept so the reference's yml files (experiments/cfgs/*.yml) merge without a
KeyError (lib/detect/config.py:233-257 rejects unknown keys and mismatched types; so does this one).
"""
from __future__ import annotations

import os
import os.path as osp
import pickle

import numpy as np


class EasyDict(dict):
    """dict with attribute access; nested dicts are converted on assignment."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


edict = EasyDict

__C = edict()
cfg = __C

# ---- training options (config.py:36-97): unused by the hot path, kept for yml compatibility ----
__C.TRAIN = edict(dict(
    SCALES=(600,), MAX_SIZE=1000, IMS_PER_BATCH=2, BATCH_SIZE=128, FG_FRACTION=0.25, AZ_POS_FRACTION=0.5,
    FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.1, USE_FLIPPED=True, BBOX_REG=True, BBOX_THRESH=0.5,
    SNAPSHOT_ITERS=10000, USE_CACHE=False, SNAPSHOT_INFIX='', USE_PREFETCH=False,
    ADDREGIONS=[[0, 0, 1, 1], [0, 0, 0.8, 0.8], [0, 0.2, 0.8, 1], [0.2, 0, 1, 0.8], [0.2, 0.2, 1, 1]],
    UN_NORMALIZE=False, NUM_PROPOSALS=2000, ANCHORS_PER_IMG=20))

# ---- testing options (config.py:100-129) ----
__C.TEST = edict(dict(SCALES=(600,), MAX_SIZE=1000, NMS=0.5, SVM=False, BBOX_REG=True, DISPLAY=False,
                      NUM_PROPOSALS=300))

# ---- search options (config.py:135-192) ----
__C.SEAR = edict(dict(
    SUBREGION=[[0, 0, 1, 1], [-0.5, 0, 0.5, 1], [0.5, 0, 1.5, 1], [0, -0.5, 1, 0.5], [0, 0.5, 1, 1.5],
               [0, 0, 0.5, 1], [0.5, 0, 1, 1], [0, 0, 1, 0.5], [0, 0.5, 1, 1], [0.25, 0, 0.75, 1], [0, 0.25, 1, 0.75]],
    ZOOM_ERR_PROB=0.3, TRAIN_REP=8, ADJ_THRESH=0.1, EMB_OBJ_THRESH=0.5, EMB_REG_THRESH=0.25, SCALE_ADJ_CONF=False,
    Tc=0.05, FIXED_PROPOSAL_NUM=True, APPEND_BOXES=False, MIN_SIDE=10, BATCH_SIZE=10000,
    AZ_CONV=['conv5_3'], FRCNN_CONV=['conv5_3']))
__C.SEAR.NUM_SUBREG = len(__C.SEAR.SUBREGION)
__C.SEAR.APPEND_TEMP = np.transpose(np.array([[[0, 0, 1, 1], [-0.25, 0, 1, 1], [0, 0, 1.25, 1], [0, -0.25, 1, 1],
                                               [0, 0, 1, 1.25], [-0.125, -0.125, 1.125, 1.125],
                                               [0.125, 0.125, 0.875, 0.875]]]), axes=[0, 2, 1])
# NOTE: SEAR.Tz and SEAR.NUM_PROPOSALS exist only after cfg_set_mode (config.py:272-280).

# ---- misc (config.py:195-219) ----
__C.DEDUP_BOXES = 1. / 16.
__C.PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])
__C.RNG_SEED = 3
__C.EPS = 1e-14
__C.ROOT_DIR = osp.abspath(osp.join(osp.dirname(__file__), '..', '..'))


def get_output_dir(imdb, net):
    """<ROOT>/output/<EXP_DIR>/<imdb.name>[/<net.name>]  (config.py:221-231)."""
    path = osp.abspath(osp.join(__C.ROOT_DIR, 'output', __C.EXP_DIR, imdb.name))
    return path if net is None else osp.join(path, net.name)


def _merge_a_into_b(a, b):
    """Recursive, type-checked merge; unknown keys raise KeyError, mismatched types ValueError (config.py:233-257)."""
    if type(a) is not edict:
        return
    for k, v in a.items():
        if k not in b:
            raise KeyError('{} is not a valid config key'.format(k))
        if k == 'PIXEL_MEANS':
            v = np.array(v)
        if isinstance(b[k], tuple) and isinstance(v, list):
            v = tuple(v)                       # yaml has no tuple type
        if type(b[k]) is not type(v):
            raise ValueError('Type mismatch ({} vs. {}) for config key: {}'.format(type(b[k]), type(v), k))
        if type(v) is edict:
            try:
                _merge_a_into_b(a[k], b[k])
            except Exception:
                print('Error under config key: {}'.format(k))
                raise
        else:
            b[k] = v


def cfg_from_file(filename):
    """Load a yml config file and merge it into the defaults (config.py:259-265)."""
    import yaml
    with open(filename, 'r') as f:
        yaml_cfg = edict(yaml.safe_load(f))
    _merge_a_into_b(yaml_cfg, __C)


def cfg_set_mode(mode, thresh=None):
    """'Train': Tz = 0, 2000 proposals; 'Test': Tz = thresh (required), 300 proposals (config.py:272-280)."""
    if mode == 'Train':
        __C.SEAR.Tz = 0.0
        __C.SEAR.NUM_PROPOSALS = __C.TRAIN.NUM_PROPOSALS
    elif mode == 'Test':
        assert thresh is not None, 'testing Tz is not set!'
        __C.SEAR.Tz = thresh
        __C.SEAR.NUM_PROPOSALS = __C.TEST.NUM_PROPOSALS


def cfg_load_thresh(filename):
    """Zoom threshold written by tune_thresh as a pickle (config.py:282-287)."""
    with open(filename, 'rb') as f:
        return pickle.load(f)


def cfg_set_path(exp_dir):
    __C.EXP_DIR = 'default' if exp_dir is None else exp_dir
