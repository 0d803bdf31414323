"""Device-resident AZ-Net adaptive search for a batch of same-sized images.

This is the B200 restatement of `im_propose` (lib/detect/test.py:346-414) + `_az_forward`
(:189-257): all levels run back to back on one CUDA stream, region counts never leave HBM, and
the only host<->device traffic is the feature maps going in and the proposal lists coming out.

Per level (k = 1 .. K-1), five C-ABI calls (include/aznet_b200.h):
    azn_roi_pool_fwd   ROI max-pool of the level's unique ROIs over the cached conv5_3 maps
    azn_fc_forward x3  int6(+ReLU) -> [int7_1 | int7_2](+ReLU) -> [adj_score | adj_bbox | zoom_score](+sigmoid)
    azn_search_level   decode/clip/unwrap, zoom select, divide_region, _sift_dup, next-level dedup + pack
followed by azn_select_proposals (top-N / Tc).

Data layout in HBM: feature maps NHWC bf16 [n_img, H, W, C]; pooled rows [M, 7*7*C] bf16 with K index
(ph*7+pw)*C + c, so the `int6` weight columns are permuted once at load time from Caffe's
c*49 + ph*7 + pw; activations bf16, head outputs f32 [M, 64]
(cols 0-10 adj_prob, 11-54 adj_bbox, 55 zoom_prob).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import ops


# The head's three output layers: "mma" = azn_az_heads_forward (small mma.sync kernel), "gemm" = the persistent tcgen05
# kernel with the AZ-head epilogue (azn_fc_forward); same arithmetic, A/B switch for benchmarks.
HEADS_KERNEL = "mma"
# Levels of the search whose ROI pool asks for the shared-memory-staged kernel (per-image buckets, azn_roi_pool_fwd_ex) instead
# of the direct L2-fed one: an A/B switch for benchmarks (bench.py --pool-staged-levels); same bits either way.
POOL_STAGED_LEVELS = ()
# From this level on the ROI pool runs with one CTA per ROI (azn_roi_pool_fwd_ex kernel_choice 3): the deep levels hold many
# small ROIs, whose seven bin rows share map rows through L1 when they run on one SM (level 5: 0.0325 -> 0.0254 ms, level 4:
# 0.0242 -> 0.0227 per 64 images; level 3 is 9 % slower that way and level 2, nine large ROIs per image, 2.6 x).  0: never.
POOL_PER_ROI_FROM_LEVEL = 4


def im_scale_for(im_h, im_w, scales=(600,), max_size=1000):
    """Image -> network-input scale, _get_image_blob (lib/detect/test.py:40-52), first TEST.SCALES entry."""
    size_min, size_max = min(im_h, im_w), max(im_h, im_w)
    s = float(scales[0]) / float(size_min)
    if np.round(s * size_max) > max_size:
        s = float(max_size) / float(size_max)
    return s


def search_depth(im_h, im_w, min_side):
    """K of lib/detect/test.py:363-368 (`side/MIN_SIDE` is Python-2 floor division for an int MIN_SIDE)."""
    side = int(min(im_h, im_w))
    q = side // min_side if isinstance(min_side, (int, np.integer)) else side / min_side
    return int(np.log2(q) + 1.0)


class AZHeadWeights:
    """AZ-Net fc weights prepared for the tensor-core path (bf16, fused, permuted), on one device."""

    def __init__(self, weights: dict, device, pooled=7):
        g = lambda n: (torch.from_numpy(np.ascontiguousarray(weights[n][0])), torch.from_numpy(np.ascontiguousarray(weights[n][1])))
        w6, b6 = g("int6")
        self.h6, k6 = w6.shape
        self.pooled = pooled
        self.C = k6 // (pooled * pooled)
        # Caffe K order c*49 + p  ->  pooled-row order p*C + c
        w6 = w6.to(device).view(self.h6, self.C, pooled * pooled).transpose(1, 2).contiguous().view(self.h6, k6)
        self.w6 = w6.to(torch.bfloat16).contiguous()
        self.b6 = b6.to(device).float().contiguous()
        w71, b71 = g("int7_1")
        w72, b72 = g("int7_2")
        self.h71, self.h72 = w71.shape[0], w72.shape[0]
        self.w7 = torch.cat([w71, w72], 0).to(device).to(torch.bfloat16).contiguous()       # one N = h71+h72 GEMM
        self.b7 = torch.cat([b71, b72], 0).to(device).float().contiguous()
        ws, bs = g("adj_score")
        wb, bb = g("adj_bbox")
        wz, bz = g("zoom_score")
        self.nsub = ws.shape[0]
        nh = 5 * self.nsub + 1
        wh = torch.zeros((nh, self.h71 + self.h72), dtype=torch.float32)
        wh[:self.nsub, :self.h71] = ws
        wh[self.nsub:5 * self.nsub, :self.h71] = wb
        wh[5 * self.nsub, self.h71:] = wz[0]
        self.wh = wh.to(device).to(torch.bfloat16).contiguous()                              # block-diagonal head
        self.bh = torch.cat([bs, bb, bz], 0).to(device).float().contiguous()
        self.n_head = nh
        self.ld_head = (nh + 7) // 8 * 8

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.w6, self.w7, self.wh))


class SearchEngine:
    """Buffers + launch sequence for the batched search.  One instance per (n_img, image shape, config)."""

    def __init__(self, head: AZHeadWeights, n_img, im_h, im_w, *, scales=(600,), max_size=1000, min_side=10, tz=0.5,
                 tc=0.05, fixed_num=True, num_proposals=300, batch_size=10000, dedup=1. / 16., eps=1e-14,
                 spatial_scale=0.0625, device=None, merge_root=True, tune=False):
        L.require_device()
        self.head = head
        self.dev = device or head.w6.device
        self.n_img, self.im_h, self.im_w = int(n_img), int(im_h), int(im_w)
        self.min_side, self.tz, self.tc = min_side, float(tz), float(tc)
        self.fixed_num, self.num_proposals = bool(fixed_num), int(num_proposals)
        self.spatial_scale = float(spatial_scale)
        self.scale = im_scale_for(im_h, im_w, scales, max_size)
        self.K = search_depth(im_h, im_w, min_side)
        # tune=True is the diagnostic search of lib/detect/tune.py:256-316: K levels (`for k in xrange(K)`), the
        # root's own zoom score against Tz = 0 at the first level, and the anchor history Bhis
        self.tune = bool(tune)
        self.n_levels = self.K if self.tune else max(self.K - 1, 0)
        # the root is always zoomed (test.py:383-384), so level 2 does not depend on level 1's head outputs:
        # both levels go through ONE pass of the heads (one stream of the 216 MB of weights less per step)
        self.merge_root = bool(merge_root) and self.n_levels >= 2 and not self.tune
        dev, i32, f64 = self.dev, torch.int32, torch.float64
        # capacity plan: full-zoom cascade of divide_region run with the product kernel itself
        self.level_sizes = self._plan_levels()
        capR = max(self.level_sizes + [1])
        root_children = self._root_children()
        capC = max(8 * capR, root_children) + 64
        capP = head.nsub * max(sum(self.level_sizes), 1)
        self.cap_out = capP if not fixed_num else max(min(self.num_proposals, capP), 1)
        capC = max(capC, self.cap_out)
        self.capR, self.capC, self.capP = capR, capC, capP
        n = self.n_img
        z = lambda *s, dt=i32: torch.zeros(s, dtype=dt, device=dev)
        self.im_h_d = torch.full((n,), self.im_h, dtype=i32, device=dev)
        self.im_w_d = torch.full((n,), self.im_w, dtype=i32, device=dev)
        self.im_scale_d = torch.full((n,), self.scale, dtype=f64, device=dev)
        self.regions = [z(n, capR, 4, dt=f64), z(n, capR, 4, dt=f64)]
        self.n_regions = [z(n), z(n)]
        self.inv, self.rep, self.n_uniq, self.img_off = z(n, capR), z(n, capR), z(n), z(n + 1)
        self.rois = z(n * capR + n, 5, dt=torch.float32)
        self.m_total = z(1)
        self.children, self.hashes, self.flags = z(n, capC, 4, dt=f64), z(n, capC, dt=torch.int64), z(n, capC)
        self.props, self.prop_scores = z(n, capP, 4, dt=f64), z(n, capP, dt=torch.float32)
        self.n_props, self.n_eval, self.depth, self.status = z(n), z(n), z(n), z(1)
        self.out_boxes = z(n, self.cap_out, 4, dt=f64)
        self.out_scores = z(n, self.cap_out, dt=torch.float32)
        self.out_count = z(n)
        m_cap = n * capR
        self.m_cap_level = [n * s for s in self.level_sizes]
        if self.merge_root:
            self.m_cap_level[1] += n                       # the merged pass: root rows + level-2 rows
        m_cap = max([m_cap] + self.m_cap_level)
        k6 = head.w6.shape[1]
        self.pool5 = torch.empty((m_cap, k6), dtype=torch.bfloat16, device=dev)
        self.h6 = torch.empty((m_cap, head.h6), dtype=torch.bfloat16, device=dev)
        self.h7 = torch.empty((m_cap, head.h71 + head.h72), dtype=torch.bfloat16, device=dev)
        self.heads = torch.zeros((m_cap, head.ld_head), dtype=torch.float32, device=dev)
        self.cap_hist = max(sum(self.level_sizes), 1) if self.tune else 0
        if self.tune:
            self.hist_regions, self.hist_zoom = z(n, self.cap_hist, 4, dt=f64), z(n, self.cap_hist, dt=torch.float32)
            self.n_history = z(n)
        self.collector = None                # dist.ProposalCollector: the final lists are also copied into its next slot
        self._st = L.SearchState()
        self._cur = 0
        self.launches = 0
        self.profile = False                 # bench.py: CUDA events around every launch group of the level loop
        self.prof_events = []
        self.m_hist = z(8192)
        self._m_idx = 0
        self._fill_state(batch_size, dedup, eps)

    # ---- planning ---------------------------------------------------------------------------
    def _root_children(self):
        w, h = float(self.im_w), float(self.im_h)
        lmin, lmax = min(w, h), max(w, h)
        return 3 * int(lmax / (lmin / 2)) - 1

    def _plan_levels(self):
        sizes = []
        cur = torch.tensor([[0.0, 0.0, self.im_w - 1.0, self.im_h - 1.0]], dtype=torch.float64, device=self.dev)
        for k in range(1, self.n_levels + 1):
            sizes.append(int(cur.shape[0]))
            if k == self.n_levels:
                break
            out, cnt = ops.divide_region(cur, float(self.min_side))
            c = int(cnt.item())
            if c < 0:
                raise RuntimeError("divide_region scratch overflow while planning capacities")
            cur = out[:c].contiguous()
        return sizes

    def _fill_state(self, batch_size, dedup, eps):
        st, p = self._st, (lambda t: t.data_ptr())
        st.n_img, st.cap_regions, st.cap_children, st.cap_props = self.n_img, self.capR, self.capC, self.capP
        st.nsub, st.chunk = self.head.nsub, int(batch_size)
        st.im_h, st.im_w, st.im_scale = p(self.im_h_d), p(self.im_w_d), p(self.im_scale_d)
        st.tz, st.min_side, st.eps, st.dedup = self.tz, float(self.min_side), float(eps), float(dedup)
        st.inv, st.rep, st.n_uniq, st.img_off = p(self.inv), p(self.rep), p(self.n_uniq), p(self.img_off)
        st.rois, st.m_total = p(self.rois), p(self.m_total)
        st.children, st.hashes, st.flags = p(self.children), p(self.hashes), p(self.flags)
        st.props, st.prop_scores, st.n_props = p(self.props), p(self.prop_scores), p(self.n_props)
        st.n_eval, st.depth, st.status = p(self.n_eval), p(self.depth), p(self.status)
        if self.tune:
            st.hist_regions, st.hist_zoom, st.n_history = p(self.hist_regions), p(self.hist_zoom), p(self.n_history)
            st.cap_history = self.cap_hist
        self._point(0)

    def _point(self, cur):
        self._cur = cur
        st = self._st
        st.regions, st.n_regions = self.regions[cur].data_ptr(), self.n_regions[cur].data_ptr()
        st.next_regions, st.next_n_regions = self.regions[1 - cur].data_ptr(), self.n_regions[1 - cur].data_ptr()

    # ---- the level loop ---------------------------------------------------------------------
    def begin(self):
        self._point(0)
        L.check(L.lib().azn_search_init(C.byref(self._st), ops._stream()), "azn_search_init")
        self.launches += 1

    def run_heads(self, conv_nhwc: torch.Tensor, level: int):
        """ROI pool + the three fused fc GEMMs for the unique ROIs of `level` (1-based)."""
        hd = self.head
        mc = self.m_cap_level[level - 1]
        midx = None
        if self.profile and self._m_idx < self.m_hist.numel():
            midx = self._m_idx
            self.m_hist[midx:midx + 1].copy_(self.m_total)
            self._m_idx += 1
        ev = self._ev
        t = ev()
        pool = ops.roi_pool(conv_nhwc, self.rois[:mc], hd.pooled, self.spatial_scale, layout="NHWC",
                            n_rois=self.m_total, out=self.pool5[:mc].view(mc, hd.pooled, hd.pooled, hd.C),
                            staged=True if level in POOL_STAGED_LEVELS else None,
                            per_roi=bool(POOL_PER_ROI_FROM_LEVEL) and level >= POOL_PER_ROI_FROM_LEVEL and level not in POOL_STAGED_LEVELS)
        a = pool.view(mc, -1)
        t = ev(level, "roi_pool", t, midx)
        ops.fc_forward(a, hd.w6, hd.b6, L.ACT_RELU, m_live=self.m_total, out=self.h6[:mc])
        t = ev(level, "int6", t, midx)
        ops.fc_forward(self.h6[:mc], hd.w7, hd.b7, L.ACT_RELU, m_live=self.m_total, out=self.h7[:mc])
        t = ev(level, "int7", t, midx)
        if HEADS_KERNEL == "mma":
            ops.az_heads(self.h7[:mc], hd.wh, hd.bh, hd.nsub, m_live=self.m_total, out=self.heads[:mc])
        else:
            ops.fc_forward(self.h7[:mc], hd.wh, hd.bh, L.ACT_AZ_HEAD, hd.nsub, m_live=self.m_total, out=self.heads[:mc])
        ev(level, "heads", t, midx)
        self.launches += 1 + 3 * 2

    def _ev(self, level=None, name=None, prev=None, midx=None):
        if not self.profile:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if prev is not None:
            self.prof_events.append((level, name, prev, e, midx))
        return e

    def prof_summary(self):
        """Mean per-launch-group device time by (level, stage) with the live row counts, from the CUDA events
        recorded while `profile` was on; int6 TFLOP/s and ROI-pool GB/s from the ALGORITHMIC work."""
        torch.cuda.synchronize()
        hd = self.head
        m_hist = self.m_hist.cpu().numpy()
        acc = {}
        for level, name, e0, e1, midx in self.prof_events:
            a = acc.setdefault((level, name), [0.0, 0, 0.0])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
            a[2] += float(m_hist[midx]) if midx is not None else 0.0
        k6, n6 = hd.w6.shape[1], hd.w6.shape[0]
        levels = []
        for (level, name), (ms, cnt, msum) in sorted(acc.items()):
            ms, m = ms / cnt, msum / cnt
            row = {"level": level, "stage": name, "ms": round(ms, 4), "m": round(m, 1)}
            if name == "int6":
                row["tflops"] = 2.0 * m * n6 * k6 / (ms * 1e-3) / 1e12
                row["weight_stream_gbs"] = (hd.w6.numel() * 2 + m * k6 * 2) / (ms * 1e-3) / 1e9
            if name == "roi_pool":
                row["gbs"] = (m * (20 + k6 * 2)) / (ms * 1e-3) / 1e9
            levels.append(row)
        int6 = [r for r in levels if r["stage"] == "int6"]
        top = max(int6, key=lambda r: r["m"]) if int6 else {"tflops": 0.0, "m": 0, "ms": 0.0}
        self._m_idx = 0
        return {"levels": levels, "int6_deepest": top}

    def begin_merged(self):
        """Levels 1+2 share one pass of the heads: root ROIs in rows [0, n_img), level-2 ROIs behind them."""
        self._point(0)
        L.check(L.lib().azn_search_root(C.byref(self._st), ops._stream()), "azn_search_root")
        self.launches += 3

    def search_level(self, level: int, root_props: bool = False):
        hd, ld = self.head, self.head.ld_head
        flags = L.LEVEL_ROOT_PROPS if root_props else (L.LEVEL_LAST if level == self.n_levels else 0)
        swap = not flags or root_props
        if self.tune:
            flags |= L.LEVEL_TUNE
        base = self.heads.data_ptr()
        L.check(L.lib().azn_search_level(C.byref(self._st), base + 4 * 5 * hd.nsub, ld, base, ld, base + 4 * hd.nsub, ld,
                                         level, flags, ops._stream()), "azn_search_level")
        self.launches += 2 if swap and not root_props else 1
        if swap:
            self._point(1 - self._cur)                 # after the root's predictions, level 2 becomes current


    def select(self):
        mode = 0 if self.fixed_num else 1
        L.check(L.lib().azn_select_proposals(C.byref(self._st), mode, self.num_proposals, self.tc,
                                             self.out_boxes.data_ptr(), self.out_scores.data_ptr(),
                                             self.out_count.data_ptr(), self.cap_out, ops._stream()), "azn_select_proposals")
        self.launches += 1
        if self.collector is not None:
            self.collector.device_add(self.out_boxes, self.out_scores, self.out_count)
            self.launches += 1

    def propose(self, conv_nhwc: torch.Tensor):
        """Run the whole search on resident NHWC bf16 maps [n_img, H, W, C].  Asynchronous; results are
        in out_boxes / out_scores / out_count / n_eval / depth (device tensors)."""
        assert conv_nhwc.dtype == torch.bfloat16 and conv_nhwc.shape[0] == self.n_img and conv_nhwc.is_contiguous()
        first = 1
        if self.merge_root:
            self.begin_merged()
            self.run_heads(conv_nhwc, 2)
            self.search_level(1, root_props=True)
            self.search_level(2)
            first = 3
        else:
            self.begin()
        for k in range(first, self.n_levels + 1):
            self.run_heads(conv_nhwc, k)
            self.search_level(k)
        self.select()

    def capture(self, conv_nhwc: torch.Tensor, pre=None):
        """Capture `pre(); propose(conv_nhwc)` into a CUDA graph.  The launch sequence of the search is static
        (every count lives on the device, every kernel is launched for the capacity), so one graph replays the
        whole level loop with no per-launch host cost.  Returns (graph, launches per replay)."""
        assert not self.profile, "event profiling and graph capture are exclusive"
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):                      # warm-up on the capture stream: workspaces, attributes
            if pre is not None:
                pre()
            self.propose(conv_nhwc)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        before = self.launches
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            if pre is not None:
                pre()
            self.propose(conv_nhwc)
        return g, self.launches - before

    def history(self):
        """Anchor history per image, f64 [n_i, 5] = (region, zoom score): the `Bhis` of lib/detect/tune.py:298."""
        assert self.tune, "the anchor history is recorded by SearchEngine(tune=True)"
        torch.cuda.current_stream().synchronize()
        cnt = self.n_history.cpu().numpy()
        reg, zoom = self.hist_regions.cpu().numpy(), self.hist_zoom.cpu().numpy()
        return [np.hstack((reg[i, :cnt[i]], zoom[i, :cnt[i], None].astype(np.float64))) for i in range(self.n_img)]

    def results(self):
        """Synchronise and fetch (boxes list of f64 [n_i,4], scores list, n_eval, depth) to the host."""
        torch.cuda.current_stream().synchronize()
        st = int(self.status.item())
        if st != 0:
            raise RuntimeError("search capacity overflow on device (status %d)" % st)
        cnt = self.out_count.cpu().numpy()
        boxes = self.out_boxes.cpu().numpy()
        scores = self.out_scores.cpu().numpy()
        return ([boxes[i, :cnt[i]].copy() for i in range(self.n_img)], [scores[i, :cnt[i]].copy() for i in range(self.n_img)],
                self.n_eval.cpu().numpy(), self.depth.cpu().numpy())
