// oracle/caffe_shim/caffe/shim.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Stand-in for the parts of the Caffe framework that the reference's layer SOURCES need in order
// to compile on their own: Caffe itself cannot be built in the authoring container (no glog, gflags,
// boost, protobuf compiler, BLAS, ...).  oracle/build_ref.py compiles the
// UNMODIFIED files
//     caffe-fast-rcnn/src/caffe/layers/{roi_pooling,grn,sigmoid,softmax,relu,inner_product,pooling}_layer.cpp
// (read where they lie under /root/reference) against this directory into
// oracle/_ref/libcaffe_layers_ref.so, so that Forward_cpu of every layer on AZ-Net's search path is
// the reference's own code and only the plumbing underneath (Blob storage, parameter messages, the
// BLAS calls of math_functions) is ours.
//
// What is written here is the minimum the layer sources name: a dense host-only Blob, parameter
// structs with the accessors of the generated protobuf classes, a Layer base class, the glog
// macros (a failed CHECK throws caffe::CheckFailed instead of aborting the process, so a test can
// observe it) and the math_functions the Forward passes call.  BLAS: the reference links an
// un-vendored library (ATLAS / MKL / OpenBLAS, Makefile.config); the gemm / gemv here follow the
// NETLIB reference implementation's summation order (one float accumulation per output element, k
// ascending), which is one of the orders the reference may legitimately have.
#ifndef AZN_ORACLE_CAFFE_SHIM_HPP_
#define AZN_ORACLE_CAFFE_SHIM_HPP_

#include <algorithm>
#include <cmath>
// The real include chain of every layer source reaches <math.h> (layer.hpp -> blob.hpp -> syncedmem.hpp ->
// util/math_functions.hpp -> util/mkl_alternate.hpp:13).  It matters: with a libstdc++ whose <math.h> exports the
// std:: overloads (gcc >= 6) the unqualified `exp(-x)` of sigmoid_layer.cpp:12 resolves to exp(float); with the 2015
// toolchains it resolved to exp(double).  The two differ by at most 1 ulp of the float result; the reference's own
// test accepts 4 (test_neuron_layer.cpp:202-217).
#include <math.h>
#include <cstddef>
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef CPU_ONLY
#define CPU_ONLY 1
#endif

namespace caffe {

using std::shared_ptr;
using std::string;
using std::vector;

// ---- glog ---------------------------------------------------------------------------------
struct CheckFailed : std::runtime_error {
  explicit CheckFailed(const std::string& m) : std::runtime_error(m) {}
};

class NullStream {
 public:
  template <typename T> NullStream& operator<<(const T&) { return *this; }
};

class FailStream {
 public:
  FailStream(const char* what, const char* file, int line) { os_ << file << ":" << line << " Check failed: " << what << " "; }
  template <typename T> FailStream& operator<<(const T& v) { os_ << v; return *this; }
  [[noreturn]] ~FailStream() noexcept(false) { throw CheckFailed(os_.str()); }
 private:
  std::ostringstream os_;
};

struct Voidify { void operator&(const NullStream&) {} void operator&(const FailStream&) {} };

}  // namespace caffe

#define LOG(severity) ::caffe::NullStream()
#define DLOG(severity) ::caffe::NullStream()
#define CHECK(cond) (cond) ? (void)0 : ::caffe::Voidify() & ::caffe::FailStream(#cond, __FILE__, __LINE__)
#define CHECK_OP_(a, b, op) CHECK((a) op (b))
#define CHECK_EQ(a, b) CHECK_OP_(a, b, ==)
#define CHECK_NE(a, b) CHECK_OP_(a, b, !=)
#define CHECK_LT(a, b) CHECK_OP_(a, b, <)
#define CHECK_LE(a, b) CHECK_OP_(a, b, <=)
#define CHECK_GT(a, b) CHECK_OP_(a, b, >)
#define CHECK_GE(a, b) CHECK_OP_(a, b, >=)
#define DCHECK(c) CHECK(c)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define NOT_IMPLEMENTED CHECK(false) << "Not Implemented Yet"
// the sources end with these; instantiation is done explicitly by oracle/ref_caffe_wrap.cpp
#define STUB_GPU(classname)
#define STUB_GPU_FORWARD(classname, funcname)
#define STUB_GPU_BACKWARD(classname, funcname)
#define INSTANTIATE_CLASS(classname)
#define REGISTER_LAYER_CLASS(type)

namespace caffe {

// ---- Blob: dense row-major host array of up to N axes; may alias caller memory ---------------
template <typename Dtype>
class Blob {
 public:
  Blob() : data_(nullptr), diff_(nullptr), count_(0) {}
  explicit Blob(const vector<int>& shape) : data_(nullptr), diff_(nullptr), count_(0) { Reshape(shape); }
  Blob(int n, int c, int h, int w) : data_(nullptr), diff_(nullptr), count_(0) { Reshape(n, c, h, w); }

  void Reshape(int n, int c, int h, int w) {
    vector<int> s(4); s[0] = n; s[1] = c; s[2] = h; s[3] = w; Reshape(s);
  }
  void Reshape(const vector<int>& shape) {
    shape_ = shape;
    count_ = 1;
    for (size_t i = 0; i < shape.size(); ++i) { CHECK_GE(shape[i], 0); count_ *= shape[i]; }
    if (!aliased_ && (size_t)count_ > own_.size()) { own_.resize(count_); }
    if (!aliased_) data_ = own_.data();
  }
  void ReshapeLike(const Blob& o) { Reshape(o.shape()); }
  // alias caller-owned memory (what Blob::set_cpu_data does in Caffe)
  void set_cpu_data(Dtype* p) { data_ = p; aliased_ = true; }

  const vector<int>& shape() const { return shape_; }
  int num_axes() const { return (int)shape_.size(); }
  int CanonicalAxisIndex(int axis) const {
    CHECK_GE(axis, -num_axes()); CHECK_LT(axis, num_axes());
    return axis < 0 ? axis + num_axes() : axis;
  }
  int shape(int i) const { return shape_[CanonicalAxisIndex(i)]; }
  int count() const { return count_; }
  int count(int a, int b) const { int c = 1; for (int i = a; i < b; ++i) c *= shape_[i]; return c; }
  int count(int a) const { return count(a, num_axes()); }
  int LegacyShape(int i) const { return i < num_axes() ? shape_[i] : 1; }
  int num() const { return LegacyShape(0); }
  int channels() const { return LegacyShape(1); }
  int height() const { return LegacyShape(2); }
  int width() const { return LegacyShape(3); }
  int offset(int n, int c = 0, int h = 0, int w = 0) const {
    return ((n * channels() + c) * height() + h) * width() + w;
  }
  const Dtype* cpu_data() const { return data_; }
  Dtype* mutable_cpu_data() { return data_; }
  const Dtype* cpu_diff() const { return diff_store(); }
  Dtype* mutable_cpu_diff() { return diff_store(); }

 private:
  Dtype* diff_store() const {
    if (dstore_.size() < (size_t)count_) dstore_.resize(count_);
    return dstore_.data();
  }
  vector<int> shape_;
  vector<Dtype> own_;
  mutable vector<Dtype> dstore_;
  Dtype* data_;
  Dtype* diff_;
  int count_;
  bool aliased_ = false;
};

// ---- parameter messages (the accessors the sources call on the protobuf classes) -------------
struct FillerParameter {};
struct ROIPoolingParameter {
  int pooled_h_ = 0, pooled_w_ = 0; float spatial_scale_ = 1.f;
  int pooled_h() const { return pooled_h_; }
  int pooled_w() const { return pooled_w_; }
  float spatial_scale() const { return spatial_scale_; }
};
struct SoftmaxParameter { int axis_ = 1; int axis() const { return axis_; } };
struct ReLUParameter { float negative_slope_ = 0.f; float negative_slope() const { return negative_slope_; } };
struct InnerProductParameter {
  int num_output_ = 0; bool bias_term_ = true; int axis_ = 1; FillerParameter wf_, bf_;
  int num_output() const { return num_output_; }
  bool bias_term() const { return bias_term_; }
  int axis() const { return axis_; }
  const FillerParameter& weight_filler() const { return wf_; }
  const FillerParameter& bias_filler() const { return bf_; }
};
enum PoolingParameter_PoolMethod {
  PoolingParameter_PoolMethod_MAX = 0, PoolingParameter_PoolMethod_AVE = 1, PoolingParameter_PoolMethod_STOCHASTIC = 2
};
struct PoolingParameter {
  PoolingParameter_PoolMethod pool_ = PoolingParameter_PoolMethod_MAX;
  int kernel_size_ = 0, stride_ = 1, pad_ = 0;
  bool global_pooling() const { return false; }
  bool has_kernel_size() const { return true; }
  bool has_kernel_h() const { return false; }
  bool has_kernel_w() const { return false; }
  bool has_pad() const { return true; }
  bool has_pad_h() const { return false; }
  bool has_pad_w() const { return false; }
  bool has_stride() const { return true; }
  bool has_stride_h() const { return false; }
  bool has_stride_w() const { return false; }
  int kernel_size() const { return kernel_size_; }
  int kernel_h() const { return kernel_size_; }
  int kernel_w() const { return kernel_size_; }
  int pad() const { return pad_; }
  int pad_h() const { return pad_; }
  int pad_w() const { return pad_; }
  int stride() const { return stride_; }
  int stride_h() const { return stride_; }
  int stride_w() const { return stride_; }
  PoolingParameter_PoolMethod pool() const { return pool_; }
};
struct LayerParameter {
  ROIPoolingParameter roi_; SoftmaxParameter softmax_; ReLUParameter relu_; InnerProductParameter ip_; PoolingParameter pool_;
  const ROIPoolingParameter& roi_pooling_param() const { return roi_; }
  const SoftmaxParameter& softmax_param() const { return softmax_; }
  const ReLUParameter& relu_param() const { return relu_; }
  const InnerProductParameter& inner_product_param() const { return ip_; }
  const PoolingParameter& pooling_param() const { return pool_; }
};

template <typename Dtype>
class Filler {
 public:
  void Fill(Blob<Dtype>* b) { std::fill(b->mutable_cpu_data(), b->mutable_cpu_data() + b->count(), Dtype(0)); }
};
template <typename Dtype>
Filler<Dtype>* GetFiller(const FillerParameter&) { return new Filler<Dtype>(); }

// ---- Layer base ------------------------------------------------------------------------------
template <typename Dtype>
class Layer {
 public:
  explicit Layer(const LayerParameter& p) : layer_param_(p) {}
  virtual ~Layer() {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {}
  virtual void Reshape(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) = 0;
  vector<shared_ptr<Blob<Dtype> > >& blobs() { return blobs_; }
  // public so that the C wrapper can drive the layer the way Net::Forward does
  void SetUp(const vector<Blob<Dtype>*>& b, const vector<Blob<Dtype>*>& t) { LayerSetUp(b, t); Reshape(b, t); }
  void Forward(const vector<Blob<Dtype>*>& b, const vector<Blob<Dtype>*>& t) { Forward_cpu(b, t); }

 protected:
  virtual void Forward_cpu(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) = 0;
  virtual void Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {}
  LayerParameter layer_param_;
  vector<shared_ptr<Blob<Dtype> > > blobs_;
  vector<bool> param_propagate_down_;
};

template <typename Dtype>
class NeuronLayer : public Layer<Dtype> {
 public:
  explicit NeuronLayer(const LayerParameter& p) : Layer<Dtype>(p) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& b, const vector<Blob<Dtype>*>& t) { t[0]->ReshapeLike(*b[0]); }
};

// ---- layer class declarations: the members the .cpp files define / touch ---------------------
#define AZN_LAYER_METHODS_                                                                      \
  virtual void Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top); \
  virtual void Backward_cpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, \
                            const vector<Blob<Dtype>*>& bottom);

template <typename Dtype>
class ROIPoolingLayer : public Layer<Dtype> {
 public:
  explicit ROIPoolingLayer(const LayerParameter& p) : Layer<Dtype>(p) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  const Blob<int>& max_idx() const { return max_idx_; }
 protected:
  AZN_LAYER_METHODS_
  int channels_, height_, width_, pooled_height_, pooled_width_;
  Dtype spatial_scale_;
  Blob<int> max_idx_;
};

template <typename Dtype>
class GRNLayer : public Layer<Dtype> {
 public:
  explicit GRNLayer(const LayerParameter& p) : Layer<Dtype>(p) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
 protected:
  AZN_LAYER_METHODS_
  Blob<Dtype> sum_multiplier_, square_, norm_, temp_dot_;
};

template <typename Dtype>
class SigmoidLayer : public NeuronLayer<Dtype> {
 public:
  explicit SigmoidLayer(const LayerParameter& p) : NeuronLayer<Dtype>(p) {}
 protected:
  AZN_LAYER_METHODS_
};

template <typename Dtype>
class ReLULayer : public NeuronLayer<Dtype> {
 public:
  explicit ReLULayer(const LayerParameter& p) : NeuronLayer<Dtype>(p) {}
 protected:
  AZN_LAYER_METHODS_
};

template <typename Dtype>
class SoftmaxLayer : public Layer<Dtype> {
 public:
  explicit SoftmaxLayer(const LayerParameter& p) : Layer<Dtype>(p) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
 protected:
  AZN_LAYER_METHODS_
  int outer_num_, inner_num_, softmax_axis_;
  Blob<Dtype> sum_multiplier_, scale_;
};

template <typename Dtype>
class InnerProductLayer : public Layer<Dtype> {
 public:
  explicit InnerProductLayer(const LayerParameter& p) : Layer<Dtype>(p) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
 protected:
  AZN_LAYER_METHODS_
  int M_, K_, N_;
  bool bias_term_;
  Blob<Dtype> bias_multiplier_;
};

template <typename Dtype>
class PoolingLayer : public Layer<Dtype> {
 public:
  explicit PoolingLayer(const LayerParameter& p) : Layer<Dtype>(p) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
 protected:
  AZN_LAYER_METHODS_
  int kernel_h_, kernel_w_, stride_h_, stride_w_, pad_h_, pad_w_;
  int channels_, height_, width_, pooled_height_, pooled_width_;
  bool global_pooling_;
  Blob<Dtype> rand_idx_;
  Blob<int> max_idx_;
};

// ---- math_functions (caffe-fast-rcnn/src/caffe/util/math_functions.cpp) -----------------------
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112 };

template <typename Dtype> inline void caffe_set(const int n, const Dtype a, Dtype* y) { for (int i = 0; i < n; ++i) y[i] = a; }
template <typename Dtype> inline void caffe_copy(const int n, const Dtype* x, Dtype* y) {
  if (x != y) std::memcpy(y, x, sizeof(Dtype) * (size_t)n);
}
template <typename Dtype> inline void caffe_sqr(const int n, const Dtype* a, Dtype* y) { for (int i = 0; i < n; ++i) y[i] = a[i] * a[i]; }
template <typename Dtype> inline void caffe_exp(const int n, const Dtype* a, Dtype* y) { for (int i = 0; i < n; ++i) y[i] = std::exp(a[i]); }
template <typename Dtype> inline void caffe_powx(const int n, const Dtype* a, const Dtype b, Dtype* y) {
  for (int i = 0; i < n; ++i) y[i] = std::pow(a[i], b);
}
template <typename Dtype> inline void caffe_div(const int n, const Dtype* a, const Dtype* b, Dtype* y) { for (int i = 0; i < n; ++i) y[i] = a[i] / b[i]; }
template <typename Dtype> inline void caffe_mul(const int n, const Dtype* a, const Dtype* b, Dtype* y) { for (int i = 0; i < n; ++i) y[i] = a[i] * b[i]; }
template <typename Dtype> inline void caffe_axpy(const int n, const Dtype alpha, const Dtype* x, Dtype* y) { for (int i = 0; i < n; ++i) y[i] += alpha * x[i]; }
template <typename Dtype> inline Dtype caffe_cpu_strided_dot(const int n, const Dtype* x, const int incx, const Dtype* y, const int incy) {
  Dtype s = 0; for (int i = 0; i < n; ++i) s += x[(size_t)i * incx] * y[(size_t)i * incy]; return s;
}
// C = alpha * op(A) * op(B) + beta * C, row-major (math_functions.cpp:13-21 maps this onto cblas_sgemm)
template <typename Dtype>
inline void caffe_cpu_gemm(const CBLAS_TRANSPOSE TransA, const CBLAS_TRANSPOSE TransB, const int M, const int N, const int K,
                           const Dtype alpha, const Dtype* A, const Dtype* B, const Dtype beta, Dtype* C) {
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      Dtype acc = 0;
      for (int k = 0; k < K; ++k) {
        const Dtype a = TransA == CblasNoTrans ? A[(size_t)i * K + k] : A[(size_t)k * M + i];
        const Dtype b = TransB == CblasNoTrans ? B[(size_t)k * N + j] : B[(size_t)j * K + k];
        acc += a * b;
      }
      Dtype& c = C[(size_t)i * N + j];
      c = beta == Dtype(0) ? alpha * acc : alpha * acc + beta * c;
    }
}
// y = alpha * op(A) * x + beta * y, A row-major M x N
template <typename Dtype>
inline void caffe_cpu_gemv(const CBLAS_TRANSPOSE TransA, const int M, const int N, const Dtype alpha, const Dtype* A, const Dtype* x,
                           const Dtype beta, Dtype* y) {
  const int rows = TransA == CblasNoTrans ? M : N, cols = TransA == CblasNoTrans ? N : M;
  for (int i = 0; i < rows; ++i) {
    Dtype acc = 0;
    for (int k = 0; k < cols; ++k) acc += (TransA == CblasNoTrans ? A[(size_t)i * N + k] : A[(size_t)k * N + i]) * x[k];
    y[i] = beta == Dtype(0) ? alpha * acc : alpha * acc + beta * y[i];
  }
}

}  // namespace caffe

#endif  // AZN_ORACLE_CAFFE_SHIM_HPP_
