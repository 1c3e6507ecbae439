// TEST ORACLE — not part of the product.
//
// Thin C harness around the reference's OWN CUDA kernels.  `make -C oracle ref`
// compiles /root/reference/src/kernels/tensor_operators.cu (+ tensors/*.cu)
// where they lie and links them with this file into
// oracle/_ref/libmarian_ref_kernels.so.  Nothing of the reference is copied:
// this file only wraps raw device pointers into the reference's marian::Tensor
// and forwards to its functions (src/kernels/tensor_operators.h:19-397), so the
// sm_100a kernels of this repo can be compared kernel-vs-kernel with the real
// thing on the GPU box, and the reference's kernels can be timed there
// ("ref-GPU" column of the op table).
//
// The reference launches on the legacy default stream; every wrapper ends with
// a device synchronisation.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <vector>

#include "common/logging.h"
#include "kernels/tensor_operators.h"
#include "functional/functional.h"

// the reference defines this in common/logging.cpp, which drags in Boost.ProgramOptions
std::shared_ptr<spdlog::logger> stderrLogger(const std::string& name, const std::string& pattern, const std::vector<std::string>&, bool) {
  auto logger = spdlog::stderr_logger_mt(name);
  logger->set_pattern(pattern);
  return logger;
}

using namespace marian;

struct ref_tensor {
  float* data;
  int rank;
  int shape[4];
};

static Tensor W(const ref_tensor& t) {
  if(!t.data)
    return nullptr;
  Shape s;
  s.resize(t.rank);
  for(int i = 0; i < t.rank; ++i)
    s.set(i, t.shape[i]);
  auto mem = New<MemoryPiece>((uint8_t*)t.data, (size_t)s.elements() * sizeof(float));
  return Tensor(new TensorBase(mem, s, 0));
}
static std::vector<Tensor> WV(const ref_tensor* ts, int n) {
  std::vector<Tensor> v;
  for(int i = 0; i < n; ++i)
    v.push_back(W(ts[i]));
  return v;
}
static cublasHandle_t handle() {
  static cublasHandle_t h = nullptr;
  if(!h)
    cublasCreate(&h);
  return h;
}
static int done() {
  cudaError_t rc = cudaDeviceSynchronize();
  return rc == cudaSuccess ? 0 : (int)rc;
}

extern "C" {

int ref_softmax(ref_tensor out, ref_tensor in, const ref_tensor* mask) {
  Softmax(W(out), W(in), mask ? W(*mask) : nullptr);
  return done();
}
int ref_logsoftmax(ref_tensor out, ref_tensor in) {
  LogSoftmax(W(out), W(in));
  return done();
}
int ref_softmax_grad(ref_tensor grad, ref_tensor adj, ref_tensor val) {
  SoftmaxGrad(W(grad), W(adj), W(val));
  return done();
}
int ref_logsoftmax_grad(ref_tensor grad, ref_tensor adj, ref_tensor val) {
  LogSoftmaxGrad(W(grad), W(adj), W(val));
  return done();
}
int ref_cross_entropy_pick(ref_tensor out, ref_tensor in, ref_tensor pick) {
  CrossEntropyPick(W(out), W(in), W(pick));
  return done();
}
int ref_cross_entropy_pick_backward(ref_tensor out, ref_tensor adj, ref_tensor in, ref_tensor pick) {
  CrossEntropyPickBackward(W(out), W(adj), W(in), W(pick));
  return done();
}
int ref_layer_norm(ref_tensor out, ref_tensor in, ref_tensor gamma, const ref_tensor* beta, float eps) {
  LayerNormalization(W(out), W(in), W(gamma), beta ? W(*beta) : nullptr, eps);
  return done();
}
int ref_layer_norm_grad(ref_tensor gx, ref_tensor gg, const ref_tensor* gb, ref_tensor adj, ref_tensor y, ref_tensor x, ref_tensor gamma, const ref_tensor* beta, float eps) {
  LayerNormalizationGrad(W(gx), W(gg), gb ? W(*gb) : nullptr, W(adj), W(y), W(x), W(gamma), beta ? W(*beta) : nullptr, eps);
  return done();
}
int ref_prod(ref_tensor C, ref_tensor A, ref_tensor B, int tA, int tB, float beta, float scalar) {
  Prod(handle(), W(C), W(A), W(B), tA, tB, beta, scalar);
  return done();
}
int ref_prod_batched(ref_tensor C, ref_tensor A, ref_tensor B, int tA, int tB, float beta, float scalar) {
  ProdBatched(handle(), W(C), W(A), W(B), tA, tB, beta, scalar);
  return done();
}
int ref_gru_fast_forward(ref_tensor out, const ref_tensor* inputs, int n, int final) {
  GRUFastForward(W(out), WV(inputs, n), final != 0);
  return done();
}
int ref_gru_fast_backward(const ref_tensor* outputs, const ref_tensor* inputs, int n, ref_tensor adj, int final) {
  GRUFastBackward(WV(outputs, 4), WV(inputs, n), W(adj), final != 0);
  return done();
}
int ref_lstm_cell_forward(ref_tensor out, const ref_tensor* inputs, int n) {
  LSTMCellForward(W(out), WV(inputs, n));
  return done();
}
int ref_lstm_output_forward(ref_tensor out, const ref_tensor* inputs, int n) {
  LSTMOutputForward(W(out), WV(inputs, n));
  return done();
}
int ref_lstm_cell_backward(const ref_tensor* outputs, const ref_tensor* inputs, int n, ref_tensor adj) {
  LSTMCellBackward(WV(outputs, 4), WV(inputs, n), W(adj));
  return done();
}
int ref_lstm_output_backward(const ref_tensor* outputs, const ref_tensor* inputs, int n, ref_tensor adj) {
  LSTMOutputBackward(WV(outputs, 4), WV(inputs, n), W(adj));
  return done();
}
int ref_att(ref_tensor out, ref_tensor va, ref_tensor context, ref_tensor state) {
  Att(W(out), W(va), W(context), W(state));
  return done();
}
int ref_att_back(ref_tensor gva, ref_tensor gc, ref_tensor gs, ref_tensor va, ref_tensor context, ref_tensor state, ref_tensor adj) {
  AttBack(W(gva), W(gc), W(gs), W(va), W(context), W(state), W(adj));
  return done();
}
int ref_highway_forward(ref_tensor out, ref_tensor in1, ref_tensor in2, ref_tensor t) {
  HighwayForward(W(out), W(in1), W(in2), W(t));
  return done();
}
int ref_highway_backward(ref_tensor o1, ref_tensor o2, ref_tensor ot, ref_tensor in1, ref_tensor in2, ref_tensor t, ref_tensor adj) {
  HighwayBackward(W(o1), W(o2), W(ot), W(in1), W(in2), W(t), W(adj));
  return done();
}
int ref_transpose_nd(ref_tensor out, ref_tensor in, const int* axes) {
  TransposeND(W(out), W(in), std::vector<int>(axes, axes + in.rank));
  return done();
}
int ref_concatenate(ref_tensor out, const ref_tensor* ins, int n, int axis) {
  Concatenate(W(out), WV(ins, n), axis < 0 ? out.rank + axis : axis);
  return done();
}
int ref_copy_rows(ref_tensor out, ref_tensor in, const size_t* host_indices, size_t n) {
  CopyRows(W(out), W(in), std::vector<size_t>(host_indices, host_indices + n));
  return done();
}
int ref_paste_rows(ref_tensor out, ref_tensor in, const size_t* host_indices, size_t n) {
  PasteRows(W(out), W(in), std::vector<size_t>(host_indices, host_indices + n));
  return done();
}
int ref_shift(ref_tensor out, ref_tensor in, const int* shift, int invert) {
  Shape s;
  s.resize(in.rank);
  for(int i = 0; i < in.rank; ++i)
    s.set(i, shift[i]);
  Shift(W(out), W(in), s, invert != 0);
  return done();
}
int ref_l2norm(ref_tensor in, float* result) {
  *result = L2Norm(W(in));
  return done();
}

// Element / Add with the functors the graph nodes use (same names as mrn_element / mrn_add)
int ref_element(const char* functor, ref_tensor out, const ref_tensor* ins, int n_in, float c) {
  using namespace functional;
  std::string f = functor;
  Tensor o = W(out);
  auto in = WV(ins, n_in);
  if(f == "plus") Element(_1 = _2 + _3, o, in[0], in[1]);
  else if(f == "minus") Element(_1 = _2 - _3, o, in[0], in[1]);
  else if(f == "mult") Element(_1 = _2 * _3, o, in[0], in[1]);
  else if(f == "div") Element(_1 = _2 / _3, o, in[0], in[1]);
  else if(f == "tanh3") Element(_1 = tanh(_2 + _3 + _4), o, in[0], in[1], in[2]);
  else if(f == "swish") Element(_1 = _2 * logit(_2), o, in[0]);
  else if(f == "logit") Element(_1 = logit(_2), o, in[0]);
  else if(f == "relu") Element(_1 = ReLU(_2), o, in[0]);
  else if(f == "scale") Element(_1 = c * _2, o, in[0]);
  else if(f == "shift") Element(_1 = _2 + c, o, in[0]);
  else if(f == "neg") Element(_1 = -_2, o, in[0]);
  else if(f == "exp") Element(_1 = exp(_2), o, in[0]);
  else if(f == "square") Element(_1 = _2 * _2, o, in[0]);
  else return -1;
  return done();
}
int ref_add(const char* functor, float scale, ref_tensor out, const ref_tensor* ins, int n_in, float c) {
  using namespace functional;
  std::string f = functor;
  Tensor o = W(out);
  auto in = WV(ins, n_in);
  if(f == "id") Add(_1, scale, o, in[0]);
  else if(f == "neg") Add(-_1, scale, o, in[0]);
  else if(f == "mult") Add(_1 * _2, scale, o, in[0], in[1]);
  else if(f == "tanh_grad") Add(_1 * (1.0f - (_2 * _2)), scale, o, in[0], in[1]);
  else if(f == "swish_grad") Add(_1 * (_3 + logit(_2) * (1.f - _3)), scale, o, in[0], in[1], in[2]);
  else return -1;
  return done();
}

// Adam exactly as the reference issues it: three Element passes (optimizers/optimizers.cu:43-73)
// preceded by Norm::clip (optimizers/clippers.cu:12-17).
int ref_adam_step(ref_tensor params, ref_tensor grads, ref_tensor mt, ref_tensor vt, float eta, float beta1, float beta2, float eps, int t, float clip_norm) {
  using namespace functional;
  Tensor p = W(params), g = W(grads), m = W(mt), v = W(vt);
  if(clip_norm > 0) {
    float l2Norm = L2Norm(g);
    if(l2Norm >= clip_norm)
      Element(_1 = (clip_norm / l2Norm) * _1, g);
  }
  float denom1 = 1 - std::pow(beta1, (size_t)t);
  float denom2 = 1 - std::pow(beta2, (size_t)t);
  Element(_1 = (beta1 * _1) + ((1 - beta1) * _2), m, g);
  Element(_1 = (beta2 * _1) + ((1 - beta2) * (_2 * _2)), v, g);
  Element(_1 -= eta * (_2 / denom1) / (sqrt(_3 / denom2) + eps), p, m, v);
  return done();
}

}  // extern "C"
