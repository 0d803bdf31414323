/*
 * oracle/caffe_layers.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, scalar, single-threaded like the reference) of the
 * Caffe layers and the Cython NMS that sit on AZ-Net's adaptive-search hot
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (aznet_b200/) never does.
 *
 * Parity status:
 *   - azo_roi_pool_fwd      pinned BIT FOR BIT (values and argmax) to Forward_cpu of the
 *                           reference's own unmodified roi_pooling_layer.cpp, compiled against
 *                           oracle/caffe_shim into oracle/_ref/libcaffe_layers_ref.so
 *                           (tests/test_oracle.py, tests/golden/caffe_layers.npz).  The reference
 *                           itself holds no forward test (only a GPU gradient check,
 *                           caffe-fast-rcnn/src/caffe/test/test_roi_pooling_layer.cpp:90-101).
 *                           Secondary cross-check: torchvision.ops.roi_pool (CPU).
 *   - azo_sigmoid           pinned bit for bit to the compiled sigmoid_layer.cpp (and by
 *                           test_neuron_layer.cpp:202-217 to 4 ulp).
 *   - azo_softmax           pinned bit for bit to the compiled softmax_layer.cpp (and by
 *                           test_softmax_layer.cpp:40-72 to 1e-4).
 *   - azo_nms               pinned against the reference's own lib/utils/nms.pyx
 *                           compiled here (oracle/_ref) and by tests/golden/nms_*.npz.
 *
 * All paths below are relative to /root/reference.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/*
 * ROI max pooling, forward.
 * Follows caffe-fast-rcnn/src/caffe/layers/roi_pooling_layer.cpp:46-125
 * (ROIPoolingLayer<float>::Forward_cpu).
 *   feat   f32 NCHW [n_img, C, H, W]
 *   rois   f32 [R, 5] = (batch_index, x1, y1, x2, y2) in image-pyramid pixels
 *   out    f32 [R, C, PH, PW]
 *   argmax i32 [R, C, PH, PW] or NULL  (h*W + w of the winning cell, -1 if empty)
 * Returns 0, or -1 if a batch index is out of range (the reference CHECK-fails
 * and aborts the process there, :66-67).
 */
int azo_roi_pool_fwd(const float *feat, int n_img, int C, int H, int W,
                     const float *rois, int R, int PH, int PW,
                     float spatial_scale, float *out, int32_t *argmax)
{
    const size_t plane = (size_t)H * W;
    for (int n = 0; n < R; ++n) {
        const float *roi = rois + (size_t)n * 5;
        int b = (int)roi[0];
        if (b < 0 || b >= n_img) return -1;
        /* :62-65  C round() = half away from zero, on a float product */
        int start_w = (int)round(roi[1] * spatial_scale);
        int start_h = (int)round(roi[2] * spatial_scale);
        int end_w = (int)round(roi[3] * spatial_scale);
        int end_h = (int)round(roi[4] * spatial_scale);
        /* :69-74  malformed ROIs are forced to 1x1; bin size in float */
        int roi_h = imax(end_h - start_h + 1, 1);
        int roi_w = imax(end_w - start_w + 1, 1);
        const float bin_h = (float)roi_h / (float)PH;
        const float bin_w = (float)roi_w / (float)PW;
        const float *img = feat + (size_t)b * C * plane;
        for (int c = 0; c < C; ++c) {
            const float *src = img + (size_t)c * plane;
            for (int ph = 0; ph < PH; ++ph) {
                /* :84-96 */
                int hs = (int)floorf((float)ph * bin_h);
                int he = (int)ceilf((float)(ph + 1) * bin_h);
                hs = imin(imax(hs + start_h, 0), H);
                he = imin(imax(he + start_h, 0), H);
                for (int pw = 0; pw < PW; ++pw) {
                    int ws = (int)floorf((float)pw * bin_w);
                    int we = (int)ceilf((float)(pw + 1) * bin_w);
                    ws = imin(imax(ws + start_w, 0), W);
                    we = imin(imax(we + start_w, 0), W);
                    size_t o = (((size_t)n * C + c) * PH + ph) * PW + pw;
                    /* :55 init -FLT_MAX ; :98-104 empty bin -> 0 ; :106-114 strict > */
                    float best = -FLT_MAX;
                    int best_i = -1;
                    if (he <= hs || we <= ws) best = 0.f;
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w) {
                            float v = src[(size_t)h * W + w];
                            if (v > best) { best = v; best_i = h * W + w; }
                        }
                    out[o] = best;
                    if (argmax) argmax[o] = best_i;
                }
            }
        }
    }
    return 0;
}

/*
 * Sigmoid, caffe-fast-rcnn/src/caffe/layers/sigmoid_layer.cpp:11-13:
 *   `1. / (1. + exp(-x))` with Dtype=float -- exp evaluated in float, the add
 *   and the divide in double (the literals are double), rounded to float on
 *   return.
 */
void azo_sigmoid(const float *x, float *y, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        float e = expf(-x[i]);
        y[i] = (float)(1. / (1. + (double)e));
    }
}

/* ReLU, caffe-fast-rcnn/src/caffe/layers/relu_layer.cpp:16-19 (negative_slope 0). */
void azo_relu(float *x, size_t n)
{
    for (size_t i = 0; i < n; ++i) x[i] = x[i] > 0.f ? x[i] : 0.f;
}

/*
 * Row softmax over `ch` channels, caffe-fast-rcnn/src/caffe/layers/softmax_layer.cpp:28-60
 * with inner_num_ == 1 (InnerProduct output): max-subtract, exp, sum, divide.
 */
void azo_softmax(const float *x, float *y, size_t rows, int ch)
{
    for (size_t r = 0; r < rows; ++r) {
        const float *xi = x + r * ch;
        float *yi = y + r * ch;
        float m = xi[0];
        for (int j = 1; j < ch; ++j) m = xi[j] > m ? xi[j] : m;
        float s = 0.f;
        for (int j = 0; j < ch; ++j) { yi[j] = expf(xi[j] - m); s += yi[j]; }
        for (int j = 0; j < ch; ++j) yi[j] = yi[j] / s;
    }
}

/*
 * Greedy NMS, lib/utils/nms.pyx:17-68.
 *   dets   f32 [n, 5] (x1, y1, x2, y2, score), row stride `stride` floats
 *   order  i64 [n]    indices sorted by score descending (the caller performs
 *                     `scores.argsort()[::-1]`, :25, so that the tie rule of the
 *                     sort stays outside this function)
 *   thresh double     (`np.float thresh` is a C double, :17; the float overlap
 *                     is promoted for the compare, :65)
 *   keep   i64 [n]    out: kept indices in score order; returns their number.
 * All box arithmetic is float32, one rounding per operation (:24, :57-64).
 */
#if defined(__GNUC__)
__attribute__((optimize("fp-contract=off")))
#endif
int64_t azo_nms(const float *dets, int64_t stride, int64_t n,
                const int64_t *order, double thresh, int64_t *keep)
{
    float *area = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    uint8_t *dead = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int64_t i = 0; i < n; ++i) {
        const float *d = dets + i * stride;
        volatile float w = d[2] - d[0];
        volatile float h = d[3] - d[1];
        w = w + 1.f;
        h = h + 1.f;
        area[i] = w * h;                                     /* :24 */
    }
    int64_t nk = 0;
    for (int64_t a = 0; a < n; ++a) {
        int64_t i = order[a];
        if (dead[i]) continue;
        keep[nk++] = i;
        const float *di = dets + i * stride;
        const float ix1 = di[0], iy1 = di[1], ix2 = di[2], iy2 = di[3];
        const float ia = area[i];
        for (int64_t b = a + 1; b < n; ++b) {
            int64_t j = order[b];
            if (dead[j]) continue;
            const float *dj = dets + j * stride;
            float xx1 = ix1 >= dj[0] ? ix1 : dj[0];          /* :13-14 max is `a if a >= b else b` */
            float yy1 = iy1 >= dj[1] ? iy1 : dj[1];
            float xx2 = ix2 <= dj[2] ? ix2 : dj[2];          /* :16-17 */
            float yy2 = iy2 <= dj[3] ? iy2 : dj[3];
            volatile float w = xx2 - xx1;
            w = w + 1.f;
            volatile float h = yy2 - yy1;
            h = h + 1.f;
            float ww = 0.f >= w ? 0.f : w;                   /* :61-62 max(0.0, .) */
            float hh = 0.f >= h ? 0.f : h;
            volatile float inter = ww * hh;
            volatile float uni = ia + area[j];
            uni = uni - inter;
            float ovr = inter / uni;                         /* :64 */
            if ((double)ovr >= thresh) dead[j] = 1;          /* :65-66 */
        }
    }
    free(area);
    free(dead);
    return nk;
}
