// The adaptive-search level step of AZ-Net for sm_100a: everything the reference does on the
// host between two net.forward() calls, fused into one kernel per level with one CTA per image.
//
//   phase 1  decode + clip + un-dedup + unwrap + min-side sift      lib/detect/test.py:106-151,
//            (appends to the image's accumulated predictions Y)      :171-187, :243-251, :380-381
//   phase 2  zoom selection (root always zoomed) + divide_region     lib/detect/test.py:383-391,
//                                                                    lib/utils/div.pyx:15-76
//   phase 3  _sift_dup: hash, unique, hash-sorted order              lib/utils/div.pyx:78-88
//   phase 4  feature-space dedup of the NEXT level's ROIs            lib/detect/test.py:61-97,:212-218
//
// Ordered compaction uses warp ballots + popc and one block-level scan; the two `np.unique`
// calls are restated as an exact stable rank: element c is a representative iff no earlier
// element has its key, and its output slot is the number of representatives with a smaller
// key.  Keys are the reference's float64 dot products evaluated as exact int64 integers.
// All float64 arithmetic follows the reference's operation order with explicit _rn intrinsics
// (no FMA contraction), so regions are bit-identical to the reference's; only np.exp (float32,
// a different libm) may differ in the last bit of the decoded box widths.
#include "common.cuh"
#include "box_common.cuh"

namespace {

constexpr int LEVEL_THREADS = 512;
constexpr int AZN_LEVEL_ROOT_DIVIDE = 64;    // internal: phases 2-4 for the root only (azn_search_root)

// ---- divide_region for one region: number of children and the children themselves -------
struct DivGeom {
    int min_ind, num_long;
    double l_short, l_long;
};

__device__ __forceinline__ DivGeom div_geom(const double x1, const double y1, const double x2, const double y2) {
    DivGeom g;
    const double len0 = __dadd_rn(__dsub_rn(x2, x1), 1.0);
    const double len1 = __dadd_rn(__dsub_rn(y2, y1), 1.0);
    g.min_ind = len0 <= len1 ? 0 : 1;                        // np.argmin: first minimum
    const double lmin = g.min_ind == 0 ? len0 : len1;
    const double lmax = g.min_ind == 0 ? len1 : len0;
    g.l_short = __ddiv_rn(lmin, 2.0);
    g.num_long = (int)(unsigned int)__ddiv_rn(lmax, g.l_short);   // C cast: truncation
    g.l_long = __ddiv_rn(lmax, (double)g.num_long);
    return g;
}

__device__ __forceinline__ int div_count(const DivGeom &g) {
    return g.num_long >= 1 ? 3 * g.num_long - 1 : 0;
}

__device__ __forceinline__ void div_emit(const DivGeom &g, const double x1, const double y1, double *__restrict__ out) {
    const int nl = g.num_long;
    const double hs = __ddiv_rn(g.l_short, 2.0), hl = __ddiv_rn(g.l_long, 2.0);
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < nl; ++j) {
            const double s0 = __dmul_rn((double)k, g.l_short), s1 = __dmul_rn((double)(k + 1), g.l_short);
            const double t0 = __dmul_rn((double)j, g.l_long), t1 = __dmul_rn((double)(j + 1), g.l_long);
            double *o = out + (size_t)(k * nl + j) * 4;
            if (g.min_ind == 0) { o[0] = __dadd_rn(s0, x1); o[1] = __dadd_rn(t0, y1); o[2] = __dadd_rn(s1, x1); o[3] = __dadd_rn(t1, y1); }
            else                { o[0] = __dadd_rn(t0, x1); o[1] = __dadd_rn(s0, y1); o[2] = __dadd_rn(t1, x1); o[3] = __dadd_rn(s1, y1); }
        }
    for (int j = 0; j < nl - 1; ++j) {
        const double s0 = __dadd_rn(__dmul_rn(0.0, g.l_short), hs), s1 = __dadd_rn(__dmul_rn(1.0, g.l_short), hs);
        const double t0 = __dadd_rn(__dmul_rn((double)j, g.l_long), hl), t1 = __dadd_rn(__dmul_rn((double)(j + 1), g.l_long), hl);
        double *o = out + (size_t)(2 * nl + j) * 4;
        if (g.min_ind == 0) { o[0] = __dadd_rn(s0, x1); o[1] = __dadd_rn(t0, y1); o[2] = __dadd_rn(s1, x1); o[3] = __dadd_rn(t1, y1); }
        else                { o[0] = __dadd_rn(t0, x1); o[1] = __dadd_rn(s0, y1); o[2] = __dadd_rn(t1, x1); o[3] = __dadd_rn(s1, y1); }
    }
}

// hash of _sift_dup: np.round(regions / min_height).dot([1, 1e3, 1e6, 1e9]) as an exact integer
__device__ __forceinline__ long long sift_hash(const double *__restrict__ r, const double min_side) {
    const long long a = (long long)rint(__ddiv_rn(r[0], min_side));
    const long long b = (long long)rint(__ddiv_rn(r[1], min_side));
    const long long c = (long long)rint(__ddiv_rn(r[2], min_side));
    const long long d = (long long)rint(__ddiv_rn(r[3], min_side));
    return a + b * 1000LL + c * 1000000LL + d * 1000000000LL;
}

__global__ void search_init_kernel(azn_search_state st) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        *st.m_total = st.n_img;
        st.img_off[st.n_img] = st.n_img;
        *st.status = 0;
    }
    if (i >= st.n_img) return;
    double *r = st.regions + (size_t)i * st.cap_regions * 4;
    r[0] = 0.0; r[1] = 0.0; r[2] = (double)st.im_w[i] - 1.0; r[3] = (double)st.im_h[i] - 1.0;   // test.py:355
    st.n_regions[i] = 1;
    st.inv[(size_t)i * st.cap_regions] = 0;
    st.rep[(size_t)i * st.cap_regions] = 0;
    st.n_uniq[i] = 1;
    st.img_off[i] = i;
    float p[4];
    project_roi(r, st.im_scale[i], p);
    float *o = st.rois + (size_t)i * 5;
    o[0] = (float)i; o[1] = p[0]; o[2] = p[1]; o[3] = p[2]; o[4] = p[3];
    st.n_props[i] = 0;
    st.n_eval[i] = 0;
    st.depth[i] = 0;
    if (st.n_history) st.n_history[i] = 0;
}

__global__ void __launch_bounds__(LEVEL_THREADS)
search_level_kernel(azn_search_state st, const float *__restrict__ zoom_prob, int ld_zoom,
                    const float *__restrict__ adj_prob, int ld_prob, const float *__restrict__ adj_bbox, int ld_bbox,
                    int level, int mode) {
    pdl_enter();
    __shared__ int s_warp[33];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int capR = st.cap_regions;
    const bool last_level = mode & AZN_LEVEL_LAST, root_props = mode & AZN_LEVEL_ROOT_PROPS,
               root_divide = mode & AZN_LEVEL_ROOT_DIVIDE;
    // merged levels 1+2: the root's rows are [0, n_img) of the head outputs and its inv/rep are the identity
    // (the per-image index arrays already describe level 2 when the root's predictions are appended)
    const int nR = st.n_regions[i];
    const int row0 = root_props ? i : st.img_off[i];
    const double *regions = st.regions + (size_t)i * capR * 4;
    const int *inv = st.inv + (size_t)i * capR;
    const int *rep = st.rep + (size_t)i * capR;
    if (nR <= 0) {                                   // search already ended for this image (test.py:388-389)
        if (tid == 0 && !root_props) st.next_n_regions[i] = 0, st.n_uniq[i] = 0;
        return;
    }
    const bool tune = mode & AZN_LEVEL_TUNE;
    if (tid == 0 && !root_divide) {
        st.depth[i] = level;
        st.n_eval[i] += nR;
    }
    // anchor history: Bhis = vstack((Bhis, hstack((B, zoom)))) (lib/detect/tune.py:298)
    if (!root_divide && st.hist_regions != nullptr) {
        const int h0 = st.n_history[i];
        __syncthreads();
        double *hr = st.hist_regions + (size_t)i * st.cap_history * 4;
        float *hz = st.hist_zoom + (size_t)i * st.cap_history;
        for (int r = tid; r < nR; r += LEVEL_THREADS) {
            if (h0 + r < st.cap_history) {
                const double *b = regions + (size_t)r * 4;
                double *o = hr + (size_t)(h0 + r) * 4;
                o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3];
                hz[h0 + r] = zoom_prob[(size_t)(row0 + (root_props ? 0 : inv[r])) * ld_zoom];
            } else {
                *st.status = AZN_ERR_CAPACITY;
            }
        }
        if (tid == 0) st.n_history[i] = h0 + nR < st.cap_history ? h0 + nR : st.cap_history;
    }

    // ---------------- phase 1: adjacent predictions -> Y ------------------------------------
    if (!root_divide) {
        const int nsub = st.nsub, ncand = nR * nsub;
        const double wmax = (double)st.im_w[i] - 1.0, hmax = (double)st.im_h[i] - 1.0;
        double *props = st.props + (size_t)i * st.cap_props * 4;
        float *pscores = st.prop_scores + (size_t)i * st.cap_props;
        int base = st.n_props[i];
        for (int c0 = 0; c0 < ncand; c0 += LEVEL_THREADS) {
            const int c = c0 + tid;
            double b[4];
            float score = 0.f;
            int keep = 0;
            if (c < ncand) {
                const int r = c / nsub, s = c - r * nsub;
                const int u = root_props ? 0 : inv[r];             // un-dedup: pred[inv_index] (:246-249)
                const double *box = regions + (size_t)(root_props ? 0 : rep[u]) * 4;  // boxes = boxes[index] (:218): the representative's box
                const float *d = adj_bbox + (size_t)(row0 + u) * ld_bbox + 4 * s;
                decode_clip(box[0], box[1], box[2], box[3], d[0], d[1], d[2], d[3], st.eps, wmax, hmax, b);
                const double hh = __dadd_rn(__dsub_rn(b[3], b[1]), 1.0), ww = __dadd_rn(__dsub_rn(b[2], b[0]), 1.0);
                keep = (hh < ww ? hh : ww) >= st.min_side ? 1 : 0; // _unwrap_adj_pred (:182-185)
                score = adj_prob[(size_t)(row0 + u) * ld_prob + s];
            }
            int total;
            const int pos = base + block_excl_scan(keep, total, s_warp);
            if (keep) {
                if (pos < st.cap_props) {
                    double *o = props + (size_t)pos * 4;
                    o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3];
                    pscores[pos] = score;
                } else {
                    *st.status = AZN_ERR_CAPACITY;
                }
            }
            base += total;
        }
        if (tid == 0) st.n_props[i] = base < st.cap_props ? base : st.cap_props;
    }
    if (root_props) return;
    if (last_level) {
        if (tid == 0) st.next_n_regions[i] = 0, st.n_uniq[i] = 0;
        return;
    }

    // ---------------- phase 2: zoom selection + divide_region -------------------------------
    double *children = st.children + (size_t)i * st.cap_children * 4;
    long long *hashes = (long long *)st.hashes + (size_t)i * st.cap_children;
    int *flags = st.flags + (size_t)i * st.cap_children;
    int nC = 0;
    for (int r0 = 0; r0 < nR; r0 += LEVEL_THREADS) {
        const int r = r0 + tid;
        int cnt = 0;
        DivGeom g;
        const double *box = regions + (size_t)(r < nR ? r : 0) * 4;
        if (r < nR) {
            // the root region is always divided (:383-384), whatever the net says: level 2 does not depend on
            // level 1's outputs, which is what lets the two levels share one pass of the heads
            double z = 1.0;
            if (tune || !(level == 1 && r == 0)) z = (double)zoom_prob[(size_t)(row0 + inv[r]) * ld_zoom];
            const double tz = (tune && level == 1) ? 0.0 : st.tz;   // tune.py:278 starts at Tz = 0, then cfg.SEAR.Tz (:305)
            if (z >= tz) {                                      // indZ = where(zoom >= Tz) (:386)
                g = div_geom(box[0], box[1], box[2], box[3]);
                cnt = div_count(g);
            }
        }
        int total;
        const int off = nC + block_excl_scan(cnt, total, s_warp);
        if (cnt > 0) {
            if (off + cnt <= st.cap_children) div_emit(g, box[0], box[1], children + (size_t)off * 4);
            else *st.status = AZN_ERR_CAPACITY;
        }
        nC += total;
    }
    if (nC > st.cap_children) nC = st.cap_children;
    __syncthreads();

    // ---------------- phase 3: _sift_dup -> next level's regions ----------------------------
    for (int c = tid; c < nC; c += LEVEL_THREADS) hashes[c] = sift_hash(children + (size_t)c * 4, st.min_side);
    __syncthreads();
    mark_first(hashes, flags, nC);
    __syncthreads();
    double *next = st.next_regions + (size_t)i * capR * 4;
    int nN = 0;
    for (int c0 = 0; c0 < nC; c0 += LEVEL_THREADS) {
        const int c = c0 + tid;
        const int first = (c < nC) ? flags[c] : 0;
        if (first) {
            const int slot = unique_slot(hashes, flags, nC, hashes[c]);
            if (slot < capR) {
                const double *s = children + (size_t)c * 4;
                double *o = next + (size_t)slot * 4;
                o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3];
            } else {
                *st.status = AZN_ERR_CAPACITY;
            }
        }
        nN += __syncthreads_count(first);
    }
    if (nN > capR) nN = capR;
    __syncthreads();
    if (tid == 0) st.next_n_regions[i] = nN;

    // ---------------- phase 4: feature-space dedup of the next level ------------------------
    int *ninv = st.inv + (size_t)i * capR, *nrep = st.rep + (size_t)i * capR;
    if (st.dedup > 0.0) {
        const double scale = st.im_scale[i];
        const float fd = (float)st.dedup;
        const int chunk = st.chunk > 0 ? st.chunk : 0x7fffffff;
        for (int q = tid; q < nN; q += LEVEL_THREADS) {
            float p[4];
            project_roi(next + (size_t)q * 4, scale, p);
            // chunk id above the hash (hash < 1e15 < 2^50): unique() runs per chunk of BATCH_SIZE (:202-218)
            hashes[q] = feat_hash(p, fd) + ((long long)(q / chunk) << 50);
        }
        __syncthreads();
        mark_first(hashes, flags, nN);
        __syncthreads();
        int nU = 0;
        for (int q0 = 0; q0 < nN; q0 += LEVEL_THREADS) {
            const int q = q0 + tid;
            int first = 0;
            if (q < nN) {
                first = flags[q];
                const int slot = unique_slot(hashes, flags, nN, hashes[q]);
                ninv[q] = slot;
                if (first) nrep[slot] = q;
            }
            nU += __syncthreads_count(first);
        }
        if (tid == 0) st.n_uniq[i] = nU;
    } else {
        for (int q = tid; q < nN; q += LEVEL_THREADS) ninv[q] = q, nrep[q] = q;
        if (tid == 0) st.n_uniq[i] = nN;
    }
    (void)lane;
}

// Packs the per-image unique ROIs of the level that is about to run into one dense [M,5] blob.
// Called after the caller swapped regions <-> next_regions.
__global__ void __launch_bounds__(256) search_pack_kernel(azn_search_state st, int row_base) {
    pdl_enter();
    const int i = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_off;
    if (tid < 32) {
        int acc = 0;
        for (int j = tid; j < i; j += 32) acc += st.n_uniq[j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (tid == 0) {
            acc += row_base;                         // merged levels 1+2: rows [0, n_img) belong to the roots
            s_off = acc;
            st.img_off[i] = acc;
            if (i == st.n_img - 1) {
                st.img_off[st.n_img] = acc + st.n_uniq[i];
                *st.m_total = acc + st.n_uniq[i];
            }
        }
    }
    __syncthreads();
    const int off = s_off, nU = st.n_uniq[i];
    const double *regions = st.regions + (size_t)i * st.cap_regions * 4;
    const int *rep = st.rep + (size_t)i * st.cap_regions;
    const double scale = st.im_scale[i];
    for (int u = tid; u < nU; u += blockDim.x) {
        float p[4];
        project_roi(regions + (size_t)rep[u] * 4, scale, p);
        float *o = st.rois + (size_t)(off + u) * 5;
        o[0] = (float)i; o[1] = p[0]; o[2] = p[1]; o[3] = p[2]; o[4] = p[3];
    }
}

// ---- final selection (lib/detect/test.py:393-401) ---------------------------------------
constexpr int SEL_THREADS = 1024;
constexpr int SEL_STAGE = 2048;            // winners whose keys fit the shared-memory rank stage

__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(azn_search_state st, int mode, int num_proposals, double tc, double *__restrict__ out_boxes,
              float *__restrict__ out_scores, int32_t *__restrict__ out_count, int cap_out, int *__restrict__ scratch) {
    pdl_enter();
    __shared__ int s_warp[33];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_prefix, s_need;
    const int i = blockIdx.x, tid = threadIdx.x;
    const int n = st.n_props[i];
    const double *props = st.props + (size_t)i * st.cap_props * 4;
    const float *scores = st.prop_scores + (size_t)i * st.cap_props;
    double *ob = out_boxes + (size_t)i * cap_out * 4;
    float *os = out_scores + (size_t)i * cap_out;
    if (mode == 1) {                                            // Y[aScores >= Tc], original order
        int base = 0;
        for (int c0 = 0; c0 < n; c0 += SEL_THREADS) {
            const int c = c0 + tid;
            const int keep = (c < n && (double)scores[c] >= tc) ? 1 : 0;
            int total;
            const int pos = base + block_excl_scan(keep, total, s_warp);
            if (keep && pos < cap_out) {
                for (int k = 0; k < 4; ++k) ob[(size_t)pos * 4 + k] = props[(size_t)c * 4 + k];
                os[pos] = scores[c];
            }
            base += total;
        }
        if (tid == 0) {
            if (base > cap_out) *st.status = AZN_ERR_CAPACITY;
            out_count[i] = base < cap_out ? base : cap_out;
        }
        return;
    }
    // mode 0: the num_proposals highest scores, ties by lower index (stable argsort of -score)
    int k = num_proposals < n ? num_proposals : n;
    if (k > cap_out) { k = cap_out; if (tid == 0) *st.status = AZN_ERR_CAPACITY; }
    if (k <= 0) {
        if (tid == 0) out_count[i] = 0;
        return;
    }
    unsigned prefix = 0, need = (unsigned)k;    // find the k-th largest key by 8-bit radix select
    int *cand = scratch + (size_t)i * cap_out;   // candidate indices (index order)
    if (k < n) {
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int b = tid; b < 256; b += SEL_THREADS) s_hist[b] = 0;
            __syncthreads();
            const unsigned himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int c = tid; c < n; c += SEL_THREADS) {
                const unsigned key = score_key(scores[c]);
                if ((key & himask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid < 32) {
                // warp-parallel top-down scan of the 256 bins: lane l owns bins 255-8l .. 248-8l
                unsigned c[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = s_hist[255 - (8 * tid + j)]; sum += c[j]; }
                const unsigned incl = (unsigned)warp_incl_scan((int)sum, tid), excl = incl - sum;
                if (excl < need && need <= incl) {            // exactly one lane: the bin holding the need-th largest
                    unsigned acc = excl;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (acc + c[j] >= need) {
                            s_prefix = prefix | ((unsigned)(255 - (8 * tid + j)) << shift);
                            s_need = need - acc;
                            break;
                        }
                        acc += c[j];
                    }
                }
            }
            __syncthreads();
            prefix = s_prefix;
            need = s_need;
        }
    }
    // prefix = key of the k-th largest; `need` of the elements equal to it are taken, lowest index first
    const unsigned kth = prefix;
    int base = 0, eq_seen = 0;
    for (int c0 = 0; c0 < n; c0 += SEL_THREADS) {
        const int c = c0 + tid;
        unsigned key = 0;
        int gt = 0, eq = 0;
        if (c < n) {
            key = score_key(scores[c]);
            gt = (k >= n) ? 1 : (key > kth);
            eq = (k < n) && key == kth;
        }
        int tot_eq;
        const int eq_pos = eq_seen + block_excl_scan(eq, tot_eq, s_warp);
        const int take = gt || (eq && eq_pos < (int)need);
        int total;
        const int pos = base + block_excl_scan(take, total, s_warp);
        if (take && pos < cap_out) cand[pos] = c;
        base += total;
        eq_seen += tot_eq;
    }
    __syncthreads();
    const int m = base < cap_out ? base : cap_out;             // == k
    // exact rank of every winner among the winners (score desc, index asc).  The usual case (300 proposals) stages the
    // winners' keys and indices in shared memory: the quadratic loop then reads broadcast shared words instead of a
    // dependent pair of global loads per step (measured 26 us -> a few us for the whole kernel).
    __shared__ unsigned s_key[SEL_STAGE];
    __shared__ int s_idx[SEL_STAGE];
    if (m <= SEL_STAGE) {
        for (int a = tid; a < m; a += SEL_THREADS) {
            const int ca = cand[a];
            s_idx[a] = ca;
            s_key[a] = score_key(scores[ca]);
        }
        __syncthreads();
        for (int a = tid; a < m; a += SEL_THREADS) {
            const int ca = s_idx[a];
            const unsigned ka = s_key[a];
            int rank = 0;
#pragma unroll 4
            for (int b = 0; b < m; ++b) rank += (s_key[b] > ka || (s_key[b] == ka && s_idx[b] < ca)) ? 1 : 0;
            const double2 *src = reinterpret_cast<const double2 *>(props + (size_t)ca * 4);
            double2 *dst = reinterpret_cast<double2 *>(ob + (size_t)rank * 4);
            dst[0] = src[0];
            dst[1] = src[1];
            os[rank] = scores[ca];
        }
    } else {
        for (int a = tid; a < m; a += SEL_THREADS) {
            const int ca = cand[a];
            const unsigned ka = score_key(scores[ca]);
            int rank = 0;
            for (int b = 0; b < m; ++b) {
                const int cb = cand[b];
                const unsigned kb = score_key(scores[cb]);
                rank += (kb > ka || (kb == ka && cb < ca)) ? 1 : 0;
            }
            for (int q = 0; q < 4; ++q) ob[(size_t)rank * 4 + q] = props[(size_t)ca * 4 + q];
            os[rank] = scores[ca];
        }
    }
    if (tid == 0) out_count[i] = m;
}

// ---- stand-alone divide_region / _sift_dup (one CTA) --------------------------------------
__global__ void __launch_bounds__(LEVEL_THREADS)
divide_kernel(const double *__restrict__ regions, int n, double min_side, double *__restrict__ out,
              int32_t *__restrict__ out_count, int cap_out, int sift_only, double *__restrict__ children,
              long long *__restrict__ hashes, int *__restrict__ flags, int cap_children) {
    __shared__ int s_warp[33];
    const int tid = threadIdx.x;
    int nC = 0;
    if (sift_only) {
        for (int c = tid; c < n; c += LEVEL_THREADS)
            for (int k = 0; k < 4; ++k) children[(size_t)c * 4 + k] = regions[(size_t)c * 4 + k];
        nC = n;
    } else {
        for (int r0 = 0; r0 < n; r0 += LEVEL_THREADS) {
            const int r = r0 + tid;
            int cnt = 0;
            DivGeom g;
            const double *box = regions + (size_t)(r < n ? r : 0) * 4;
            if (r < n) { g = div_geom(box[0], box[1], box[2], box[3]); cnt = div_count(g); }
            int total;
            const int off = nC + block_excl_scan(cnt, total, s_warp);
            if (cnt > 0 && off + cnt <= cap_children) div_emit(g, box[0], box[1], children + (size_t)off * 4);
            nC += total;
        }
    }
    if (nC > cap_children) { nC = cap_children; if (tid == 0) *out_count = -1; }
    __syncthreads();
    for (int c = tid; c < nC; c += LEVEL_THREADS) hashes[c] = sift_hash(children + (size_t)c * 4, min_side);
    __syncthreads();
    mark_first(hashes, flags, nC);
    __syncthreads();
    int nN = 0;
    for (int c0 = 0; c0 < nC; c0 += LEVEL_THREADS) {
        const int c = c0 + tid;
        const int first = (c < nC) ? flags[c] : 0;
        if (first) {
            const int slot = unique_slot(hashes, flags, nC, hashes[c]);
            if (slot < cap_out)
                for (int k = 0; k < 4; ++k) out[(size_t)slot * 4 + k] = children[(size_t)c * 4 + k];
        }
        nN += __syncthreads_count(first);
    }
    if (tid == 0) *out_count = nN <= cap_out ? nN : -1;
}

__global__ void decode_kernel(const double *__restrict__ boxes, const float *__restrict__ deltas, int n, int ncol,
                              double eps, double wmax, double hmax, int clip, double *__restrict__ out) {
    const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (t >= (long)n * ncol) return;
    const int r = (int)(t / ncol), s = (int)(t - (long)r * ncol);
    const double *b = boxes + (size_t)r * 4;
    const float *d = deltas + ((size_t)r * ncol + s) * 4;
    double o[4];
    if (clip) decode_clip<true>(b[0], b[1], b[2], b[3], d[0], d[1], d[2], d[3], eps, wmax, hmax, o);
    else decode_clip<false>(b[0], b[1], b[2], b[3], d[0], d[1], d[2], d[3], eps, wmax, hmax, o);
    double *dst = out + ((size_t)r * ncol + s) * 4;
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2]; dst[3] = o[3];
}

int check_state(const azn_search_state *st) {
    AZN_REQUIRE(st != nullptr, "search: null state");
    AZN_REQUIRE(st->n_img > 0 && st->cap_regions > 0 && st->cap_children > 0 && st->cap_props > 0 && st->nsub > 0,
                "search: bad capacities n_img=%d cap_regions=%d cap_children=%d cap_props=%d nsub=%d", st->n_img,
                st->cap_regions, st->cap_children, st->cap_props, st->nsub);
    AZN_REQUIRE(st->im_h && st->im_w && st->im_scale && st->regions && st->n_regions && st->inv && st->rep &&
                    st->n_uniq && st->img_off && st->rois && st->m_total && st->next_regions && st->next_n_regions &&
                    st->children && st->hashes && st->flags && st->props && st->prop_scores && st->n_props &&
                    st->n_eval && st->depth && st->status,
                "search: null pointer in state");
    AZN_REQUIRE(st->min_side > 0, "search: min_side must be positive");
    AZN_REQUIRE(st->hist_regions == nullptr || (st->hist_zoom && st->n_history && st->cap_history > 0),
                "search: hist_regions needs hist_zoom, n_history and cap_history > 0");
    return AZN_OK;
}

}  // namespace

extern "C" int azn_search_init(const azn_search_state *st, azn_stream_t stream) {
    int rc = check_state(st);
    if (rc) return rc;
    AZN_CUDA(azn_launch_pdl(search_init_kernel, dim3((st->n_img + 127) / 128), dim3(128), 0, (cudaStream_t)stream, *st));
    return AZN_OK;
}

extern "C" int azn_search_level(const azn_search_state *st, const float *zoom_prob, int ld_zoom,
                                const float *adj_prob, int ld_prob, const float *adj_bbox, int ld_bbox,
                                int level, int flags, azn_stream_t stream) {
    int rc = check_state(st);
    if (rc) return rc;
    AZN_REQUIRE(zoom_prob && adj_prob && adj_bbox && ld_zoom >= 1 && ld_prob >= st->nsub && ld_bbox >= 4 * st->nsub,
                "azn_search_level: bad head pointers/strides");
    AZN_REQUIRE(level >= 1, "azn_search_level: level is 1-based");
    AZN_REQUIRE((flags & ~(AZN_LEVEL_LAST | AZN_LEVEL_ROOT_PROPS | AZN_LEVEL_TUNE)) == 0, "azn_search_level: unknown flags %d", flags);
    AZN_REQUIRE(!(flags & AZN_LEVEL_TUNE) || !(flags & AZN_LEVEL_ROOT_PROPS), "azn_search_level: AZN_LEVEL_TUNE excludes the merged root");
    AZN_REQUIRE(!(flags & AZN_LEVEL_ROOT_PROPS) || level == 1, "azn_search_level: AZN_LEVEL_ROOT_PROPS is for level 1");
    cudaStream_t s = (cudaStream_t)stream;
    AZN_CUDA(azn_launch_pdl(search_level_kernel, dim3(st->n_img), dim3(LEVEL_THREADS), 0, s, *st, zoom_prob, ld_zoom, adj_prob,
                            ld_prob, adj_bbox, ld_bbox, level, flags));
    if (!(flags & (AZN_LEVEL_LAST | AZN_LEVEL_ROOT_PROPS))) {
        // the next level's regions become current: swap, then pack its unique ROIs
        azn_search_state nx = *st;
        nx.regions = st->next_regions;
        nx.n_regions = st->next_n_regions;
        AZN_CUDA(azn_launch_pdl(search_pack_kernel, dim3(st->n_img), dim3(256), 0, s, nx, 0));
    }
    return AZN_OK;
}

extern "C" int azn_search_root(const azn_search_state *st, azn_stream_t stream) {
    int rc = check_state(st);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    AZN_CUDA(azn_launch_pdl(search_init_kernel, dim3((st->n_img + 127) / 128), dim3(128), 0, s, *st));
    AZN_CUDA(azn_launch_pdl(search_level_kernel, dim3(st->n_img), dim3(LEVEL_THREADS), 0, s, *st, (const float *)nullptr, 0,
                            (const float *)nullptr, 0, (const float *)nullptr, 0, 1, (int)AZN_LEVEL_ROOT_DIVIDE));
    azn_search_state nx = *st;
    nx.regions = st->next_regions;
    nx.n_regions = st->next_n_regions;
    AZN_CUDA(azn_launch_pdl(search_pack_kernel, dim3(st->n_img), dim3(256), 0, s, nx, st->n_img));
    return AZN_OK;
}

extern "C" int azn_select_proposals(const azn_search_state *st, int mode, int num_proposals, double tc,
                                    double *out_boxes, float *out_scores, int32_t *out_count, int cap_out,
                                    azn_stream_t stream) {
    int rc = check_state(st);
    if (rc) return rc;
    AZN_REQUIRE(out_boxes && out_scores && out_count && cap_out > 0, "azn_select_proposals: bad outputs");
    AZN_REQUIRE(mode == 0 || mode == 1, "azn_select_proposals: mode must be 0 (top-N) or 1 (Tc threshold)");
    AZN_REQUIRE(mode == 1 || num_proposals >= 0, "azn_select_proposals: num_proposals < 0");
    AZN_REQUIRE(cap_out <= st->cap_children, "azn_select_proposals: cap_out (%d) exceeds the flags scratch (%d)", cap_out,
                st->cap_children);
    AZN_CUDA(azn_launch_pdl(select_kernel, dim3(st->n_img), dim3(SEL_THREADS), 0, (cudaStream_t)stream, *st, mode, num_proposals,
                            tc, out_boxes, out_scores, out_count, cap_out, st->flags));
    return AZN_OK;
}

// ---- proposal lists of a run of batches, collected on the device ---------------------------------
// test_proposals appends every image's list and writes proposals.pkl once (lib/detect/test.py:508-539).  Here
// the final lists of batch `step` are copied into slot step % n_slots of a per-rank collection that is
// gathered once at the end of the job.  The step counter lives on the device (state[0]) so that the launch is
// identical every step and replays from a CUDA graph; the CTA that finishes last (ticket in state[1]) bumps it
// -- every CTA has read it by then.
__global__ void __launch_bounds__(256)
collect_kernel(const uint32_t *__restrict__ boxes, const uint32_t *__restrict__ scores, const uint32_t *__restrict__ counts,
               uint32_t *__restrict__ dst_boxes, uint32_t *__restrict__ dst_scores, uint32_t *__restrict__ dst_counts,
               long nb, long ns, long nc, int n_slots, unsigned *__restrict__ state) {
    pdl_enter();
    const unsigned step = *(volatile unsigned *)state;
    const size_t slot = step % (unsigned)n_slots;
    const long total = nb + ns + nc;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        if (i < nb) dst_boxes[slot * nb + i] = boxes[i];
        else if (i < nb + ns) dst_scores[slot * ns + (i - nb)] = scores[i - nb];
        else dst_counts[slot * nc + (i - nb - ns)] = counts[i - nb - ns];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&state[1], 1u) == gridDim.x - 1) {
            state[1] = 0u;
            state[0] = step + 1u;
        }
    }
}

extern "C" int azn_collect_proposals(const double *boxes, const float *scores, const int32_t *counts, int n_img, int cap_out,
                                     double *dst_boxes, float *dst_scores, int32_t *dst_counts, int n_slots,
                                     uint32_t *state, azn_stream_t stream) {
    AZN_REQUIRE(boxes && scores && counts && dst_boxes && dst_scores && dst_counts && state, "azn_collect_proposals: null pointer");
    AZN_REQUIRE(n_img > 0 && cap_out > 0 && n_slots > 0, "azn_collect_proposals: bad shape");
    const long nb = (long)n_img * cap_out * 8, ns = (long)n_img * cap_out, nc = n_img;      // 32-bit words
    const int grid = (int)((nb + ns + nc + 256 * 8 - 1) / (256 * 8));
    AZN_CUDA(azn_launch_pdl(collect_kernel, dim3(grid < 1 ? 1 : (grid > 592 ? 592 : grid)), dim3(256), 0, (cudaStream_t)stream,
                            (const uint32_t *)boxes, (const uint32_t *)scores, (const uint32_t *)counts, (uint32_t *)dst_boxes,
                            (uint32_t *)dst_scores, (uint32_t *)dst_counts, nb, ns, nc, n_slots, state));
    return AZN_OK;
}

extern "C" size_t azn_divide_region_scratch_bytes(int n) {
    // children of n regions: 3*num_long-1 each; num_long is unbounded for degenerate aspect ratios, so the
    // scratch is sized for 16 children per region (num_long <= 5) with a floor, and overflow is reported.
    const size_t cap = (size_t)(n > 0 ? n : 1) * 16 + 64;
    return cap * (4 * sizeof(double) + sizeof(long long) + sizeof(int)) + 256;
}

extern "C" int azn_divide_region(const double *regions, int n, double min_side, double *out, int32_t *out_count,
                                 int cap_out, int sift_only, void *scratch, size_t scratch_bytes, azn_stream_t stream) {
    AZN_REQUIRE(n >= 0 && out_count && min_side > 0, "azn_divide_region: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        AZN_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), s));
        return AZN_OK;
    }
    AZN_REQUIRE(regions && out && scratch && cap_out > 0, "azn_divide_region: null pointer");
    if (scratch_bytes < azn_divide_region_scratch_bytes(n)) {
        azn_set_error("azn_divide_region: scratch %zu < %zu bytes", scratch_bytes, azn_divide_region_scratch_bytes(n));
        return AZN_ERR_CAPACITY;
    }
    const size_t cap = (size_t)n * 16 + 64;
    double *children = (double *)scratch;
    long long *hashes = (long long *)(children + cap * 4);
    int *flags = (int *)(hashes + cap);
    divide_kernel<<<1, LEVEL_THREADS, 0, s>>>(regions, n, min_side, out, out_count, cap_out, sift_only, children, hashes,
                                             flags, (int)cap);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

extern "C" int azn_decode_boxes(const double *boxes, const float *deltas, int n, int ncol, double eps, int im_h,
                                int im_w, double *out, azn_stream_t stream) {
    AZN_REQUIRE(n >= 0 && ncol > 0, "azn_decode_boxes: bad shape");
    if (n == 0) return AZN_OK;
    AZN_REQUIRE(boxes && deltas && out, "azn_decode_boxes: null pointer");
    const long total = (long)n * ncol;
    decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(boxes, deltas, n, ncol, eps,
                                                                                   (double)im_w - 1.0, (double)im_h - 1.0,
                                                                                   (im_h > 0 && im_w > 0) ? 1 : 0, out);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}
