// oracle/ref_caffe_wrap.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points around the REFERENCE'S OWN layer sources.  The #includes below pull the unmodified
// files in from /root/reference/caffe-fast-rcnn/src/caffe/layers (the directory is given with -I by
// oracle/build_ref.py; nothing is copied into this repository); they compile against the stand-in
// framework headers of oracle/caffe_shim.  Each entry drives one layer the way caffe::Net does --
// LayerSetUp, Reshape, Forward_cpu -- over caller-owned buffers.  A failed CHECK (the reference
// aborts the process there) is reported as return code -1.
//
// Output: oracle/_ref/libcaffe_layers_ref.so (git-ignored; travels to the GPU box prebuilt).
#include <cfloat>
#include <cstdint>

#include "caffe/shim.hpp"

#include "roi_pooling_layer.cpp"    // ROIPoolingLayer<Dtype>::Forward_cpu :46-125
#include "grn_layer.cpp"            // GRNLayer<Dtype>::Forward_cpu :27-56
#include "sigmoid_layer.cpp"        // SigmoidLayer<Dtype>::Forward_cpu :11-24
#include "softmax_layer.cpp"        // SoftmaxLayer<Dtype>::Forward_cpu :28-60
#include "relu_layer.cpp"           // ReLULayer<Dtype>::Forward_cpu :10-20
#include "inner_product_layer.cpp"  // InnerProductLayer<Dtype>::Forward_cpu :80-93
#include "pooling_layer.cpp"        // PoolingLayer<Dtype>::Forward_cpu :128-229 (MAX, ceil mode :93-95)

namespace {

using caffe::Blob;
using caffe::LayerParameter;
using std::vector;

template <typename L>
int run(L& layer, vector<Blob<float>*> bottom, vector<Blob<float>*> top) {
  try {
    layer.SetUp(bottom, top);
    layer.Forward(bottom, top);
  } catch (const caffe::CheckFailed&) {
    return -1;
  }
  return 0;
}

Blob<float>* alias(Blob<float>& b, const float* p, vector<int> shape) {
  b.set_cpu_data(const_cast<float*>(p));
  b.Reshape(shape);
  return &b;
}

}  // namespace

extern "C" {

// feat f32 NCHW [n_img,C,H,W]; rois f32 [R,5]; out f32 [R,C,PH,PW]; argmax i32 [R,C,PH,PW] or NULL
int ref_roi_pool_fwd(const float* feat, int n_img, int C, int H, int W, const float* rois, int R, int PH, int PW,
                     float spatial_scale, float* out, int32_t* argmax) {
  LayerParameter p;
  p.roi_.pooled_h_ = PH; p.roi_.pooled_w_ = PW; p.roi_.spatial_scale_ = spatial_scale;
  caffe::ROIPoolingLayer<float> layer(p);
  Blob<float> b0, b1, t0;
  alias(b0, feat, {n_img, C, H, W});
  alias(b1, rois, {R, 5, 1, 1});
  t0.set_cpu_data(out);
  int rc = run(layer, {&b0, &b1}, {&t0});
  if (rc == 0 && argmax) std::memcpy(argmax, layer.max_idx().cpu_data(), sizeof(int32_t) * (size_t)R * C * PH * PW);
  return rc;
}

// x, y f32 [n,C,H,W]: y = x / sqrt(sum_c x^2) per position
int ref_grn_fwd(const float* x, int n, int C, int H, int W, float* y) {
  caffe::GRNLayer<float> layer((LayerParameter()));
  Blob<float> b0, t0;
  alias(b0, x, {n, C, H, W});
  t0.set_cpu_data(y);
  return run(layer, {&b0}, {&t0});
}

int ref_sigmoid_fwd(const float* x, float* y, int n) {
  caffe::SigmoidLayer<float> layer((LayerParameter()));
  Blob<float> b0, t0;
  alias(b0, x, {n});
  t0.set_cpu_data(y);
  return run(layer, {&b0}, {&t0});
}

int ref_relu_fwd(const float* x, float* y, int n) {
  caffe::ReLULayer<float> layer((LayerParameter()));
  Blob<float> b0, t0;
  alias(b0, x, {n});
  t0.set_cpu_data(y);
  return run(layer, {&b0}, {&t0});
}

// x, y f32 [rows, ch]; softmax over axis 1 (the prototxts' default)
int ref_softmax_fwd(const float* x, float* y, int rows, int ch) {
  caffe::SoftmaxLayer<float> layer((LayerParameter()));
  Blob<float> b0, t0;
  alias(b0, x, {rows, ch});
  t0.set_cpu_data(y);
  return run(layer, {&b0}, {&t0});
}

// x f32 [M,K] (any trailing axes flattened by the caller), w f32 [N,K], b f32 [N] or NULL, y f32 [M,N]
int ref_inner_product_fwd(const float* x, int M, int K, const float* w, const float* b, int N, float* y) {
  LayerParameter p;
  p.ip_.num_output_ = N; p.ip_.bias_term_ = b != nullptr;
  caffe::InnerProductLayer<float> layer(p);
  Blob<float> b0, t0;
  alias(b0, x, {M, K});
  t0.set_cpu_data(y);
  try {
    layer.SetUp({&b0}, {&t0});   // allocates zero-filled parameter blobs, as the fillers of a fresh net would
    std::memcpy(layer.blobs()[0]->mutable_cpu_data(), w, sizeof(float) * (size_t)N * K);
    if (b) std::memcpy(layer.blobs()[1]->mutable_cpu_data(), b, sizeof(float) * (size_t)N);
    layer.Forward({&b0}, {&t0});
  } catch (const caffe::CheckFailed&) {
    return -1;
  }
  return 0;
}

// MAX pooling, x f32 [n,C,H,W] -> y f32 [n,C,PH,PW]; writes the pooled dims the layer's Reshape computed
int ref_max_pool_fwd(const float* x, int n, int C, int H, int W, int kernel, int stride, int pad, float* y, int y_capacity,
                     int* PH, int* PW) {
  LayerParameter p;
  p.pool_.kernel_size_ = kernel; p.pool_.stride_ = stride; p.pool_.pad_ = pad;
  caffe::PoolingLayer<float> layer(p);
  Blob<float> b0, t0;
  alias(b0, x, {n, C, H, W});
  try {
    vector<Blob<float>*> bottom{&b0}, top{&t0};
    layer.SetUp(bottom, top);
    *PH = t0.height(); *PW = t0.width();
    if (t0.count() > y_capacity) return -2;
    layer.Forward(bottom, top);
    std::memcpy(y, t0.cpu_data(), sizeof(float) * (size_t)t0.count());
  } catch (const caffe::CheckFailed&) {
    return -1;
  }
  return 0;
}

}  // extern "C"
