// The memory-bound pieces of the conv5_3 backbone: image -> network-input blob, and the 2x2 max
// pooling between the conv stages.  The convolutions themselves are azn_conv3x3_forward (gemm.cu).
//
// Every activation map is a ZERO-BORDERED channels-last grid [n_img, H+2, W+2, C] bf16 (see gemm.cu): each
// kernel here writes the border of its output itself, so buffers can be reused across layers without memsets.
#include <float.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ void st16(void *p, const uint4 &v) { *reinterpret_cast<uint4 *>(p) = v; }

// ---- _get_image_blob + im_list_to_blob (lib/detect/test.py:27-59, lib/utils/blob.py:13-29) -------------------
//   im_orig = im.astype(float32) - PIXEL_MEANS     (numpy evaluates the in-place subtract in float64, stores f32)
//   cv2.resize(im_orig, fx = fy = im_scale, INTER_LINEAR)
// cv::resize, float source, INTER_LINEAR (OpenCV modules/imgproc/src/resize.cpp, 4.x):
//   scale = 1 / f (double);  for every destination index d:  s = (d + 0.5) * scale - 0.5 (double),
//   i = floor(s), a = (float)(s - i);  x: i < 0 -> (0, a = 0), i >= W-1 -> (W-1, a = 0);  y: rows i and i+1 clamped;
//   horizontal pass  t = S[i] * (1 - a) + S[i+1] * a  (float), vertical pass  d = t0 * (1 - b) + t1 * b  (float).
// One thread group of 8 lanes per destination pixel of the padded grid: lane 0 computes the three channels, all
// 8 lanes store one 16-byte piece of the pixel's Cpad-channel row (channels >= 3 are zero).
__global__ void __launch_bounds__(256)
image_blob_kernel(const uint8_t *__restrict__ im, int n_img, int H0, int W0, int Hs, int Ws, double scale_x, double scale_y,
                  double m0, double m1, double m2, __nv_bfloat16 *__restrict__ out, int Cpad, float *__restrict__ blob_f32) {
    const int vec_per_px = Cpad / 8;
    const long total = (long)n_img * (Hs + 2) * (Ws + 2) * vec_per_px;
    for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const int part = (int)(g % vec_per_px);
        const long px = g / vec_per_px;
        const int xp = (int)(px % (Ws + 2));
        const int yp = (int)((px / (Ws + 2)) % (Hs + 2));
        const int n = (int)(px / ((long)(Ws + 2) * (Hs + 2)));
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        const bool interior = xp >= 1 && xp <= Ws && yp >= 1 && yp <= Hs;
        if (part == 0 && interior) {
            const int dx = xp - 1, dy = yp - 1;
            const double xd = ((double)dx + 0.5) * scale_x - 0.5, yd = ((double)dy + 0.5) * scale_y - 0.5;
            int sx = (int)floor(xd);
            float fx = (float)(xd - (double)sx);        // the fraction is taken in double, then rounded (pinned against cv2 4.13)
            if (sx < 0) { fx = 0.f; sx = 0; }
            if (sx >= W0 - 1) { fx = 0.f; sx = W0 - 1; }
            const int sy = (int)floor(yd);
            const float fy = (float)(yd - (double)sy);
            const int y0 = min(max(sy, 0), H0 - 1), y1 = min(max(sy + 1, 0), H0 - 1);
            const int x1 = min(sx + 1, W0 - 1);
            const float a0 = __fsub_rn(1.f, fx), a1 = fx, b0 = __fsub_rn(1.f, fy), b1 = fy;
            const uint8_t *r0 = im + ((size_t)n * H0 + y0) * W0 * 3, *r1 = im + ((size_t)n * H0 + y1) * W0 * 3;
            const double mean[3] = {m0, m1, m2};
            float res[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float p00 = (float)((double)r0[sx * 3 + c] - mean[c]), p01 = (float)((double)r0[x1 * 3 + c] - mean[c]);
                const float p10 = (float)((double)r1[sx * 3 + c] - mean[c]), p11 = (float)((double)r1[x1 * 3 + c] - mean[c]);
                const float t0 = __fadd_rn(__fmul_rn(p00, a0), __fmul_rn(p01, a1));
                const float t1 = __fadd_rn(__fmul_rn(p10, a0), __fmul_rn(p11, a1));
                res[c] = __fadd_rn(__fmul_rn(t0, b0), __fmul_rn(t1, b1));
                if (blob_f32) blob_f32[(((size_t)n * 3 + c) * Hs + dy) * Ws + dx] = res[c];      // Caffe's 'data' blob, NCHW
            }
            const __nv_bfloat162 lo = __floats2bfloat162_rn(res[0], res[1]);
            const __nv_bfloat162 hi = __floats2bfloat162_rn(res[2], 0.f);
            v.x = *reinterpret_cast<const uint32_t *>(&lo);
            v.y = *reinterpret_cast<const uint32_t *>(&hi);
        }
        st16(out + (size_t)px * Cpad + part * 8, v);
    }
}

// ---- PoolingLayer MAX 2x2 stride 2 (caffe-fast-rcnn/src/caffe/layers/pooling_layer.cpp:81-95,144-168) --------
// pooled size = ceil((H - 2) / 2) + 1 (ceil mode), windows clipped to the map, strict `>` scan from -FLT_MAX.
// One thread per output pixel (of the padded grid) and 16-byte channel vector.
__device__ __forceinline__ unsigned pick_bf16x2(unsigned v, unsigned b) {
    const unsigned m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162 *>(&v), *reinterpret_cast<const __nv_bfloat162 *>(&b));
    return (v & m) | (b & ~m);
}
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const uint4 *__restrict__ in, int n_img, int H, int W, int L, uint4 *__restrict__ out, int Ho, int Wo) {
    // 32-bit indexing (the host checks the sizes) and the four window loads issued together: a window cut by the edge
    // (ceil mode, odd H or W) re-reads its last row / column -- the reduction is idempotent -- instead of looping over
    // data-dependent bounds, which kept one load in flight per thread and spent more instructions on 64-bit div / mod than
    // on the pooling (3.4 TB/s)
    const unsigned total = (unsigned)n_img * (unsigned)(Ho + 2) * (unsigned)(Wo + 2) * (unsigned)L;
    const unsigned wpo = (unsigned)(Wo + 2), hpo = (unsigned)(Ho + 2), uL = (unsigned)L;
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        const unsigned px = g / uL, v = g - px * uL;
        const unsigned row = px / wpo, xp = px - row * wpo;
        const unsigned n = row / hpo, yp = row - n * hpo;
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
        if (xp >= 1u && xp <= (unsigned)Wo && yp >= 1u && yp <= (unsigned)Ho) {
            const int hs = (int)(yp - 1u) * 2, ws = (int)(xp - 1u) * 2;
            const int h1 = min(hs + 1, H - 1), w1 = min(ws + 1, W - 1);
            const uint4 *r0 = in + ((size_t)(n * (unsigned)(H + 2) + (unsigned)hs + 1u) * (unsigned)(W + 2) + 1u) * uL + v;
            const uint4 *r1 = in + ((size_t)(n * (unsigned)(H + 2) + (unsigned)h1 + 1u) * (unsigned)(W + 2) + 1u) * uL + v;
            const uint4 q00 = __ldg(r0 + (size_t)ws * uL), q01 = __ldg(r0 + (size_t)w1 * uL);
            const uint4 q10 = __ldg(r1 + (size_t)ws * uL), q11 = __ldg(r1 + (size_t)w1 * uL);
            acc = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);       // bf16(-FLT_MAX) = -inf
            // the reference's scan order: (hs, ws), (hs, ws+1), (hs+1, ws), (hs+1, ws+1), `v > best ? v : best`
            acc.x = pick_bf16x2(q00.x, acc.x); acc.y = pick_bf16x2(q00.y, acc.y); acc.z = pick_bf16x2(q00.z, acc.z); acc.w = pick_bf16x2(q00.w, acc.w);
            acc.x = pick_bf16x2(q01.x, acc.x); acc.y = pick_bf16x2(q01.y, acc.y); acc.z = pick_bf16x2(q01.z, acc.z); acc.w = pick_bf16x2(q01.w, acc.w);
            acc.x = pick_bf16x2(q10.x, acc.x); acc.y = pick_bf16x2(q10.y, acc.y); acc.z = pick_bf16x2(q10.z, acc.z); acc.w = pick_bf16x2(q10.w, acc.w);
            acc.x = pick_bf16x2(q11.x, acc.x); acc.y = pick_bf16x2(q11.y, acc.y); acc.z = pick_bf16x2(q11.z, acc.z); acc.w = pick_bf16x2(q11.w, acc.w);
        }
        out[(size_t)g] = acc;
    }
}

// Unpadded NHWC bf16 [n, H, W, C] <-> zero-bordered grid (tests, and maps that arrive from elsewhere).
__global__ void __launch_bounds__(256)
pad_border_kernel(const uint4 *__restrict__ in, int n_img, int H, int W, int L, uint4 *__restrict__ out, int to_padded) {
    const long total = (long)n_img * (H + 2) * (W + 2) * L;
    for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const int v = (int)(g % L);
        const long px = g / L;
        const int xp = (int)(px % (W + 2));
        const int yp = (int)((px / (W + 2)) % (H + 2));
        const int n = (int)(px / ((long)(W + 2) * (H + 2)));
        const bool interior = xp >= 1 && xp <= W && yp >= 1 && yp <= H;
        const size_t flat = (((size_t)n * H + (yp - 1)) * W + (xp - 1)) * L + v;
        if (to_padded) out[(size_t)px * L + v] = interior ? __ldg(in + flat) : make_uint4(0u, 0u, 0u, 0u);
        else if (interior) out[flat] = __ldg(in + (size_t)px * L + v);
    }
}

// ---- conv1_1 as a one-tap GEMM: the 3x3 neighbourhood of every pixel gathered once ("patches") ----------------
// With Cin = 3 the implicit GEMM of azn_conv3x3_forward multiplies nine 64-channel taps of which 61 channels are zero
// padding (24.6 TFLOP/s of useful work, 0.86 ms per 16 images).  Here the 9 * Cin real values of a pixel's
// neighbourhood become ONE K = Kp row -- entry (ky*3+kx)*Cin + c = in[y+ky-1, x+kx-1, c], the K order of
// pack_conv_weight -- over the same zero-bordered grid, so the convolution is a single 64-deep k-block per pixel tile
// (azn_conv_patches_forward).  in: [n, H+2, W+2, Cs] zero-bordered (the border IS the padding), out: [n, H+2, W+2, Kp].
// Eight lanes per pixel, one 16-byte vector of the row each: coalesced 128-byte stores.
__global__ void __launch_bounds__(256)
patches3x3_kernel(const __nv_bfloat16 *__restrict__ in, int n_img, int H, int W, int Cs, int Cin, __nv_bfloat16 *__restrict__ out, int Kp) {
    const int vpp = Kp / 8;
    const long total = (long)n_img * (H + 2) * (W + 2) * vpp;
    const int nk = 9 * Cin;
    // the grid stride is a multiple of vpp (host), so a thread keeps its 16-byte part of the K row for every pixel it
    // visits: the eight (neighbour pixel, channel) offsets are computed once
    const long g0 = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const int part = (int)(g0 % vpp);
    long off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int e = part * 8 + j;
        off[j] = -1;
        if (e < nk) {
            const int tap = e / Cin, c = e - tap * Cin, ky = tap / 3, kx = tap - ky * 3;
            off[j] = ((long)(ky - 1) * (W + 2) + (kx - 1)) * Cs + c + (long)(W + 3) * Cs;      // biased by one row + one pixel: >= 0
        }
    }
    const unsigned short *src = reinterpret_cast<const unsigned short *>(in) - (long)(W + 3) * Cs;
    const bool any = off[0] >= 0;                                  // entries are dense from 0: part past 9*Cin is all zero
    for (long g = g0; g < total; g += (long)gridDim.x * blockDim.x) {
        const long px = g / vpp;
        const int xp = (int)(px % (W + 2));
        const int yp = (int)((px / (W + 2)) % (H + 2));
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (any && xp >= 1 && xp <= W && yp >= 1 && yp <= H) {
            const unsigned short *p = src + (size_t)px * Cs;
            unsigned e8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) e8[j] = off[j] >= 0 ? (unsigned)__ldg(p + off[j]) : 0u;
            v.x = e8[0] | (e8[1] << 16); v.y = e8[2] | (e8[3] << 16);
            v.z = e8[4] | (e8[5] << 16); v.w = e8[6] | (e8[7] << 16);
        }
        st16(out + (size_t)px * Kp + part * 8, v);
    }
}

// ---- conv1_1 without the patch matrix (round 2) --------------------------------------------------------------------
// The patches route writes and re-reads a 128-byte K row per pixel (786 MB per 16 images: 0.41 ms to gather + 0.37 ms for
// the one-tap GEMM) to feed a product whose real K is 27.  Here a warp takes 16 consecutive pixels of the zero-bordered
// grid, builds the two k16 A fragments of mma.sync.m16n8k16 straight from the 8-channel network input (a lane needs 8
// neighbourhood values of 2 pixels; the input is 16 bytes per pixel and lives in L1 / L2), multiplies them by the
// register-resident weight fragments (K = 32, N = 64: 16 MMAs), adds the bias, applies the ReLU, and hands the 16 x 64 tile
// through a 2 KB shared staging row so that the warp's output leaves as one contiguous run of 16-byte stores.  Border
// pixels write zeros (the next layer's padding).  Bound by the 792 MB of output per 16 images.
constexpr int C1_WARPS = 8;
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(C1_WARPS * 32, 2)
conv3x3_direct_kernel(const uint4 *__restrict__ in, int n_img, int H, int W, int Cin, const unsigned short *__restrict__ wt,
                      int Kp, const float *__restrict__ bias, uint4 *__restrict__ out, int relu) {
    // in: one 16-byte vector (8 channels) per pixel of the zero-bordered grid.  The 3x3 neighbourhoods of 16 consecutive flat
    // pixels p0 .. p0 + 15 are three runs of 18 consecutive flat pixels (p0 - 1 + (ky - 1)(W + 2) ...): 54 vectors, staged per
    // warp in shared memory; the next tile's vectors are in flight while this tile is multiplied.
    __shared__ __align__(16) unsigned short s_in[C1_WARPS][3 * 18 * 8];
    __shared__ __align__(16) unsigned short s_out[C1_WARPS][16][64 + 8];      // +8: rows 144 B apart, conflict-free fragment stores
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nk = 9 * Cin;
    // K entries of this lane: kstep u, half h (k + 8), element e: k = 16 u + 8 h + 2 t + e  ->  position in the staged runs
    int koff[2][2][2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 16 * u + 8 * h + 2 * t + e;
                koff[u][h][e] = -1;
                if (k < nk) {
                    const int tap = k / Cin, c = k - tap * Cin, ky = tap / 3, kx = tap - ky * 3;
                    koff[u][h][e] = (ky * 18 + kx) * 8 + c;           // + 8 * (pixel of the tile)
                }
            }
    // weight fragments: B[k][n] = wt[n * Kp + k]; b0 = k 2t, 2t+1 (+16 u), b1 = the same + 8; column n = 8 nt + g
    uint32_t fb[2][8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const unsigned short *wr = wt + (size_t)(8 * nt + g) * Kp;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            fb[u][nt][0] = (uint32_t)wr[16 * u + 2 * t] | ((uint32_t)wr[16 * u + 2 * t + 1] << 16);
            fb[u][nt][1] = (uint32_t)wr[16 * u + 8 + 2 * t] | ((uint32_t)wr[16 * u + 8 + 2 * t + 1] << 16);
        }
    }
    const long total = (long)n_img * (H + 2) * (W + 2);
    const long n_tiles = (total + 15) / 16;
    const long stride = (long)gridDim.x * C1_WARPS;
    // staging vector v (0 .. 53) of a tile: run ky = v / 18, pixel p0 - 1 + (ky - 1)(W + 2) + v % 18 (zero outside the buffer)
    auto fetch = [&](long tile, int v) -> uint4 {
        if (tile >= n_tiles || v >= 54) return make_uint4(0u, 0u, 0u, 0u);
        const int ky = v / 18, j = v - ky * 18;
        const long px = tile * 16 - 1 + (long)(ky - 1) * (W + 2) + j;
        return px >= 0 && px < total ? __ldg(in + px) : make_uint4(0u, 0u, 0u, 0u);
    };
    long tile = (long)blockIdx.x * C1_WARPS + warp;
    uint4 v0 = fetch(tile, lane), v1 = fetch(tile, lane + 32);
    for (; tile < n_tiles; tile += stride) {
        uint4 *stage = reinterpret_cast<uint4 *>(s_in[warp]);
        stage[lane] = v0;
        if (lane + 32 < 54) stage[lane + 32] = v1;
        __syncwarp();
        v0 = fetch(tile + stride, lane);                               // the next tile's input: in flight during this tile's math
        v1 = fetch(tile + stride, lane + 32);
        // rows g and g + 8 of the tile = pixels tile * 16 + g (+ 8)
        bool interior[2];
        uint32_t fa[2][4];
        // (row, column) of the tile's first pixel by two 32-bit divisions (the grid has < 2^31 pixels: host check); the 16
        // pixels of the tile then wrap rows by subtraction -- 64-bit div / mod per pixel cost more than the 16 MMAs
        const unsigned p0 = (unsigned)tile * 16u, wp = (unsigned)(W + 2), hp = (unsigned)(H + 2);
        const unsigned row0 = p0 / wp, x0 = p0 - row0 * wp, y0 = row0 % hp;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            unsigned xp = x0 + (unsigned)(g + 8 * r), yp = y0;
            while (xp >= wp) { xp -= wp; ++yp; }
            if (yp >= hp) yp -= hp;                                   // the next image's grid
            interior[r] = (long)p0 + g + 8 * r < total && xp >= 1u && xp <= (unsigned)W && yp >= 1u && yp <= (unsigned)H;
            const unsigned short *p = s_in[warp] + (g + 8 * r) * 8;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t lo = koff[u][h][0] >= 0 ? (uint32_t)p[koff[u][h][0]] : 0u;
                    const uint32_t hi = koff[u][h][1] >= 0 ? (uint32_t)p[koff[u][h][1]] : 0u;
                    fa[u][r + 2 * h] = lo | (hi << 16);          // a0: row g k-low, a1: row g+8 k-low, a2: row g k-high, a3: row g+8 k-high
                }
        }
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float2 bb = __ldg(reinterpret_cast<const float2 *>(bias + 8 * nt + 2 * t));
            acc[nt][0] = acc[nt][2] = bb.x;
            acc[nt][1] = acc[nt][3] = bb.y;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma_bf16_16816(acc[nt], fa[u], fb[u][nt][0], fb[u][nt][1]);
        // C fragment: c0, c1 = row g, columns 8 nt + 2t, + 1; c2, c3 = row g + 8.  Border pixels (and the tail) are zero rows.
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float c0 = acc[nt][2 * r], c1 = acc[nt][2 * r + 1];
                if (relu) { c0 = fmaxf(c0, 0.f); c1 = fmaxf(c1, 0.f); }
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(c0, c1);
                *reinterpret_cast<uint32_t *>(&s_out[warp][g + 8 * r][8 * nt + 2 * t]) = interior[r] ? *reinterpret_cast<const uint32_t *>(&b2) : 0u;
            }
        __syncwarp();
        // 16 pixels x 128 bytes = 128 vectors of 16 bytes, contiguous in the output grid
#pragma unroll
        for (int v = lane; v < 128; v += 32) {
            const int row = v >> 3, part = v & 7;
            const long px = tile * 16 + row;
            if (px < total) out[(size_t)px * 8 + part] = *reinterpret_cast<const uint4 *>(&s_out[warp][row][part * 8]);
        }
        __syncwarp();
    }
}

int grid_for(long total) {
    const long blocks = (total + 255) / 256;
    const long cap = (long)azn_num_sms() * 16;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

extern "C" int azn_image_blob(const uint8_t *images, int n_img, int H0, int W0, double im_scale, const double *pixel_means,
                              void *out_padded_nhwc, int Cpad, int Hs, int Ws, float *blob_f32, azn_stream_t stream) {
    AZN_REQUIRE(images && out_padded_nhwc && pixel_means, "azn_image_blob: null pointer");
    AZN_REQUIRE(n_img > 0 && H0 > 0 && W0 > 0 && Hs > 0 && Ws > 0 && im_scale > 0, "azn_image_blob: bad shape");
    AZN_REQUIRE(Cpad >= 8 && Cpad % 8 == 0, "azn_image_blob: Cpad=%d must be a multiple of 8 (>= 8)", Cpad);
    const long total = (long)n_img * (Hs + 2) * (Ws + 2) * (Cpad / 8);
    const double sc = 1.0 / im_scale;                     // cv::resize: scale_x = 1. / inv_scale_x with inv_scale_x = fx
    image_blob_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(images, n_img, H0, W0, Hs, Ws, sc, sc, pixel_means[0],
                                                                        pixel_means[1], pixel_means[2],
                                                                        (__nv_bfloat16 *)out_padded_nhwc, Cpad, blob_f32);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

extern "C" int azn_patches3x3(const void *in_padded, int n_img, int H, int W, int Cs, int Cin, void *out_padded, int Kp,
                              azn_stream_t stream) {
    AZN_REQUIRE(in_padded && out_padded, "azn_patches3x3: null pointer");
    AZN_REQUIRE(n_img > 0 && H > 0 && W > 0 && Cin > 0 && Cs >= Cin, "azn_patches3x3: bad shape");
    AZN_REQUIRE(Kp % 8 == 0 && Kp >= 9 * Cin, "azn_patches3x3: Kp=%d must be a multiple of 8 and >= 9*Cin=%d", Kp, 9 * Cin);
    AZN_REQUIRE(256 % (Kp / 8) == 0, "azn_patches3x3: Kp / 8 = %d must divide the block size 256", Kp / 8);
    const long total = (long)n_img * (H + 2) * (W + 2) * (Kp / 8);
    patches3x3_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)in_padded, n_img, H, W, Cs, Cin,
                                                                        (__nv_bfloat16 *)out_padded, Kp);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

extern "C" int azn_conv3x3_direct_forward(const void *in_padded, int n_img, int H, int W, int Cs, int Cin, const void *wt, int Kp,
                                          const float *bias, void *out_padded, int Cout, int relu, azn_stream_t stream) {
    AZN_REQUIRE(in_padded && wt && bias && out_padded, "azn_conv3x3_direct_forward: null pointer");
    AZN_REQUIRE(n_img > 0 && H > 0 && W > 0 && Cin > 0 && Cs >= Cin, "azn_conv3x3_direct_forward: bad shape");
    AZN_REQUIRE(9 * Cin <= 32 && Kp >= 32 && Cout == 64 && Cs == 8,
                "azn_conv3x3_direct_forward: needs 9*Cin <= 32, Kp >= 32, Cout == 64, 8 channels per input pixel (Cin=%d Kp=%d Cout=%d Cs=%d)",
                Cin, Kp, Cout, Cs);
    AZN_REQUIRE((uintptr_t)out_padded % 16 == 0 && (uintptr_t)in_padded % 16 == 0, "azn_conv3x3_direct_forward: 16-byte alignment");
    AZN_REQUIRE((double)n_img * (H + 2) * (W + 2) < 2.0e9, "azn_conv3x3_direct_forward: grid of more than 2e9 pixels");
    const long tiles = ((long)n_img * (H + 2) * (W + 2) + 15) / 16;
    const long want = (tiles + C1_WARPS - 1) / C1_WARPS, cap = (long)azn_num_sms() * 8;
    conv3x3_direct_kernel<<<(unsigned)(want < cap ? want : cap), C1_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const uint4 *)in_padded, n_img, H, W, Cin, (const unsigned short *)wt, Kp, bias, (uint4 *)out_padded, relu);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

extern "C" int azn_maxpool2x2_forward(const void *in_padded, int n_img, int H, int W, int C, void *out_padded,
                                      azn_stream_t stream) {
    AZN_REQUIRE(in_padded && out_padded, "azn_maxpool2x2_forward: null pointer");
    AZN_REQUIRE(n_img > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "azn_maxpool2x2_forward: bad shape (C must be a multiple of 8)");
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;         // ceil((H - 2) / 2) + 1
    const long total = (long)n_img * (Ho + 2) * (Wo + 2) * (C / 8);
    AZN_REQUIRE(total < (1L << 31) && (long)n_img * (H + 2) * (W + 2) < (1L << 31), "azn_maxpool2x2_forward: map too large for 32-bit indexing");
    maxpool2x2_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4 *)in_padded, n_img, H, W, C / 8,
                                                                        (uint4 *)out_padded, Ho, Wo);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}

extern "C" int azn_nhwc_border(const void *in, int n_img, int H, int W, int C, void *out, int to_padded, azn_stream_t stream) {
    AZN_REQUIRE(in && out, "azn_nhwc_border: null pointer");
    AZN_REQUIRE(n_img > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "azn_nhwc_border: bad shape (C must be a multiple of 8)");
    const long total = (long)n_img * (H + 2) * (W + 2) * (C / 8);
    pad_border_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4 *)in, n_img, H, W, C / 8, (uint4 *)out,
                                                                        to_padded ? 1 : 0);
    AZN_LAUNCH_CHECK();
    return AZN_OK;
}
