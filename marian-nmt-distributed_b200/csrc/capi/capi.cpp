// extern "C" boundary (include/marian_b200.h) over the tensor operators and the
// training-step driver.  Compiled into the product library (CUDA kernels) and,
// unchanged, into the test oracle (CPU restatement); which one answers is
// reported by mrn_backend_name().
#include "marian_b200.h"

#include <cstring>
#include <map>
#include <string>

#include "data/batch.h"
#include "data/corpus.h"
#include "kernels/tensor_operators.h"
#include "training/checkpoint.h"
#include "training/graph_group.h"
#include "training/validator.h"
#include "translator/translator.h"

using namespace marian;

namespace {

thread_local std::string g_lastError;

template <class F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch(const std::exception& e) {
    g_lastError = e.what();
    return 1;
  } catch(...) {
    g_lastError = "unknown error";
    return 2;
  }
}

Tensor wrap(const mrn_tensor& t) {
  if(!t.data)
    return nullptr;
  ABORT_IF(t.rank < 1 || t.rank > 4, "mrn_tensor rank must be 1..4");
  std::vector<int> dims(t.shape, t.shape + t.rank);
  Shape shape(dims);
  auto mem = New<MemoryPiece>((uint8_t*)t.data, (size_t)shape.elements() * sizeof(float));
  return Tensor(new TensorBase(mem, shape, device::getDevice()));
}
Tensor wrapOpt(const mrn_tensor* t) {
  return t ? wrap(*t) : nullptr;
}
std::vector<Tensor> wrapAll(const mrn_tensor* ts, int n) {
  std::vector<Tensor> v;
  for(int i = 0; i < n; ++i)
    v.push_back(wrap(ts[i]));
  return v;
}

struct Trainer {
  Ptr<Options> options;
  int device{0}, rank{0}, nranks{1};
  Ptr<SingletonGraph> single;
  Ptr<SyncGraphGroup> sync;
  Ptr<AsyncGraphGroup> async;
  Ptr<data::CorpusBatch> batch;
  Ptr<data::SyntheticCorpus> corpus;
  Ptr<data::PrefetchingBatchGenerator> textBatches;  // mrn_trainer_open_corpus
  size_t replays{0};

  GradientWorker& worker() { return single ? single->worker() : (sync ? sync->worker() : async->worker()); }
};

}  // namespace

extern "C" {

const char* mrn_last_error(void) {
  return g_lastError.c_str();
}
const char* mrn_backend_name(void) {
  return device::backendName();
}
int mrn_set_device(int d) {
  return guarded([&] { device::setDevice(d); });
}
int mrn_set_stream(void* s) {
  return guarded([&] { device::setStream(s); });
}
int mrn_synchronize(void) {
  return guarded([&] { device::synchronize(); });
}

int mrn_malloc(void** ptr, size_t bytes) {
  return guarded([&] { *ptr = device::mallocDevice(bytes); });
}
int mrn_free(void* ptr) {
  return guarded([&] { device::freeDevice(ptr); });
}
int mrn_memcpy_h2d(void* dst, const void* src, size_t bytes) {
  return guarded([&] { device::copyH2DBlocking(dst, src, bytes); });
}
int mrn_memcpy_d2h(void* dst, const void* src, size_t bytes) {
  return guarded([&] {
    void* p = device::pinnedScratch(bytes);
    device::copyD2H(p, src, bytes);
    device::synchronize();
    std::memcpy(dst, p, bytes);
  });
}
int mrn_memset_zero(void* dst, size_t bytes) {
  return guarded([&] { device::zero(dst, bytes); });
}

int mrn_gemm_create(void** handle, int dev) {
  return guarded([&] { *handle = createGemmContext(dev); });
}
int mrn_gemm_destroy(void* handle) {
  return guarded([&] { destroyGemmContext((GemmHandle)handle); });
}
int mrn_gemm_set_mode(void* handle, int mode) {
  return guarded([&] { setGemmMode((GemmHandle)handle, (GemmMode)mode); });
}

int mrn_gemm_debug_stamps(void* deviceBuffer) {
  return guarded([&] { gemmDebugStamps((unsigned long long*)deviceBuffer); });
}
int mrn_gemm_profile(int enable, double* ms, double* flops, size_t* launches) {
  return guarded([&] { gemmProfile(enable, ms, flops, launches); });
}

int mrn_prod(void* g, mrn_tensor C, mrn_tensor A, mrn_tensor B, int tA, int tB, float beta, float scalar) {
  return guarded([&] {
    gemmInvalidateCache((GemmHandle)g);
    Prod((GemmHandle)g, wrap(C), wrap(A), wrap(B), tA, tB, beta, scalar);
  });
}
int mrn_prod_batched(void* g, mrn_tensor C, mrn_tensor A, mrn_tensor B, int tA, int tB, float beta, float scalar) {
  return guarded([&] {
    gemmInvalidateCache((GemmHandle)g);
    ProdBatched((GemmHandle)g, wrap(C), wrap(A), wrap(B), tA, tB, beta, scalar);
  });
}
int mrn_prod_grouped_nt(void* g, mrn_tensor C, const mrn_tensor* As, const mrn_tensor* Bs, int n, float beta) {
  return guarded([&] {
    gemmInvalidateCache((GemmHandle)g);
    ProdGroupedNT((GemmHandle)g, wrap(C), wrapAll(As, n), wrapAll(Bs, n), beta);
  });
}
int mrn_prod_grouped_nt_sums(void* g, mrn_tensor C, const mrn_tensor* As, const mrn_tensor* Bs, int n, float beta, const mrn_tensor* col_sums) {
  return guarded([&] {
    gemmInvalidateCache((GemmHandle)g);
    ProdGroupedNT((GemmHandle)g, wrap(C), wrapAll(As, n), wrapAll(Bs, n), beta, wrapAll(col_sums, n));
    ProdFlushColumnSums((GemmHandle)g);  // sums the product queued: a pass of their own on the side stream
    device::joinSide();
  });
}
int mrn_prod_swish_grad_nt_sums(void* g, mrn_tensor C, mrn_tensor A, mrn_tensor B, mrn_tensor H, float beta, mrn_tensor col_sum) {
  return guarded([&] {
    gemmInvalidateCache((GemmHandle)g);
    ABORT_IF(!ProdSwishGradFusable((GemmHandle)g, wrap(C), wrap(A), wrap(B), wrap(H)), "mrn_prod_swish_grad_nt_sums: needs a tensor-core mode and 16-byte aligned operands");
    ProdSwishGradNT((GemmHandle)g, wrap(C), wrap(A), wrap(B), wrap(H), beta, wrap(col_sum));
    ProdFlushColumnSums((GemmHandle)g);
    device::joinSide();
  });
}
int mrn_prod_shared_a(void* g, const mrn_tensor* Cs, mrn_tensor A, const mrn_tensor* Bs, const mrn_tensor* biases, int n, int tA, float beta, int* fused) {
  return guarded([&] {
    gemmInvalidateCache((GemmHandle)g);
    auto cs = wrapAll(Cs, n), bs = wrapAll(Bs, n);
    std::vector<Tensor> bias;
    if(biases)
      bias = wrapAll(biases, n);
    Tensor a = wrap(A);
    bool done = ProdSharedA((GemmHandle)g, cs, a, bs, bias, tA != 0, beta);
    if(!done) {  // the products one by one, as callers of ProdSharedA do
      for(int i = 0; i < n; ++i) {
        if(!bias.empty() && !tA && beta == 0.f)
          ProdAffine((GemmHandle)g, cs[i], a, bs[i], bias[i]);
        else {
          ABORT_IF(!bias.empty(), "mrn_prod_shared_a: a bias needs a non-transposed, non-accumulating product");
          Prod((GemmHandle)g, cs[i], a, bs[i], tA != 0, false, beta, 1.f);
        }
      }
    }
    if(fused)
      *fused = done ? 1 : 0;
  });
}
int mrn_prod_swish_grad_nt(void* g, mrn_tensor C, mrn_tensor A, mrn_tensor B, mrn_tensor H, float beta) {
  return guarded([&] {
    gemmInvalidateCache((GemmHandle)g);
    ABORT_IF(!ProdSwishGradFusable((GemmHandle)g, wrap(C), wrap(A), wrap(B), wrap(H)), "mrn_prod_swish_grad_nt: needs the tf32 mode and 16-byte aligned operands");
    ProdSwishGradNT((GemmHandle)g, wrap(C), wrap(A), wrap(B), wrap(H), beta);
  });
}
int mrn_prod_affine(void* g, mrn_tensor C, mrn_tensor A, mrn_tensor B, mrn_tensor bias) {
  return guarded([&] {
    gemmInvalidateCache((GemmHandle)g);
    ProdAffine((GemmHandle)g, wrap(C), wrap(A), wrap(B), wrap(bias));
  });
}

int mrn_element(const char* functor, mrn_tensor out, const mrn_tensor* ins, int n_in, float c) {
  return guarded([&] {
    using namespace functional;
    std::string f = functor;
    Tensor o = wrap(out);
    auto in = wrapAll(ins, n_in);
    auto need = [&](int n) { ABORT_IF(n_in != n, "mrn_element: wrong number of inputs for functor", f); };
    if(f == "plus") { need(2); Element(_1 = _2 + _3, o, in[0], in[1]); }
    else if(f == "minus") { need(2); Element(_1 = _2 - _3, o, in[0], in[1]); }
    else if(f == "mult") { need(2); Element(_1 = _2 * _3, o, in[0], in[1]); }
    else if(f == "div") { need(2); Element(_1 = _2 / _3, o, in[0], in[1]); }
    else if(f == "tanh3") { need(3); Element(_1 = tanh(_2 + _3 + _4), o, in[0], in[1], in[2]); }
    else if(f == "tanh2") { need(2); Element(_1 = tanh(_2 + _3), o, in[0], in[1]); }
    else if(f == "tanh") { need(1); Element(_1 = tanh(_2), o, in[0]); }
    else if(f == "swish") { need(1); Element(_1 = _2 * logit(_2), o, in[0]); }
    else if(f == "logit") { need(1); Element(_1 = logit(_2), o, in[0]); }
    else if(f == "relu") { need(1); Element(_1 = ReLU(_2), o, in[0]); }
    else if(f == "scale") { need(1); Element(_1 = c * _2, o, in[0]); }
    else if(f == "shift") { need(1); Element(_1 = _2 + c, o, in[0]); }
    else if(f == "neg") { need(1); Element(_1 = -_2, o, in[0]); }
    else if(f == "exp") { need(1); Element(_1 = exp(_2), o, in[0]); }
    else if(f == "log") { need(1); Element(_1 = log(_2), o, in[0]); }
    else if(f == "sqrt") { need(1); Element(_1 = sqrt(_2 + c), o, in[0]); }
    else if(f == "square") { need(1); Element(_1 = _2 * _2, o, in[0]); }
    else if(f == "axpy_self") { need(1); Element(_1 = _1 + c * _2, o, in[0]); }
    else ABORT("mrn_element: unknown functor", f);
  });
}

int mrn_add(const char* functor, float scale, mrn_tensor out, const mrn_tensor* ins, int n_in, float c) {
  return guarded([&] {
    using namespace functional;
    std::string f = functor;
    Tensor o = wrap(out);
    auto in = wrapAll(ins, n_in);
    auto need = [&](int n) { ABORT_IF(n_in != n, "mrn_add: wrong number of inputs for functor", f); };
    if(f == "id") { need(1); Add(_1, scale, o, in[0]); }
    else if(f == "neg") { need(1); Add(-_1, scale, o, in[0]); }
    else if(f == "scale") { need(1); Add(c * _1, scale, o, in[0]); }
    else if(f == "mult") { need(2); Add(_1 * _2, scale, o, in[0], in[1]); }
    else if(f == "tanh_grad") { need(2); Add(_1 * (1.0f - (_2 * _2)), scale, o, in[0], in[1]); }
    else if(f == "logit_grad") { need(2); Add(_1 * _2 * (1.0f - _2), scale, o, in[0], in[1]); }
    else if(f == "swish_grad") { need(3); Add(_1 * (_3 + logit(_2) * (1.f - _3)), scale, o, in[0], in[1], in[2]); }
    else if(f == "relu_grad") { need(2); Add(_1 * ReLUback(_2), scale, o, in[0], in[1]); }
    else if(f == "div_grad_a") { need(2); Add(_1 * 1.0f / _2, scale, o, in[0], in[1]); }
    else if(f == "div_grad_b") { need(3); Add(-_1 * _2 / (_3 * _3), scale, o, in[0], in[1], in[2]); }
    else ABORT("mrn_add: unknown functor", f);
  });
}

int mrn_softmax(mrn_tensor out, mrn_tensor in, const mrn_tensor* mask) {
  return guarded([&] { Softmax(wrap(out), wrap(in), wrapOpt(mask)); });
}
int mrn_logsoftmax(mrn_tensor out, mrn_tensor in) {
  return guarded([&] { LogSoftmax(wrap(out), wrap(in)); });
}
int mrn_softmax_grad(mrn_tensor grad, mrn_tensor adj, mrn_tensor val) {
  return guarded([&] { SoftmaxGrad(wrap(grad), wrap(adj), wrap(val)); });
}
int mrn_logsoftmax_grad(mrn_tensor grad, mrn_tensor adj, mrn_tensor val) {
  return guarded([&] { LogSoftmaxGrad(wrap(grad), wrap(adj), wrap(val)); });
}
int mrn_cross_entropy_pick(mrn_tensor out, mrn_tensor in, mrn_tensor pick) {
  return guarded([&] { CrossEntropyPick(wrap(out), wrap(in), wrap(pick)); });
}
int mrn_cross_entropy_pick_backward(mrn_tensor out, mrn_tensor adj, mrn_tensor in, mrn_tensor pick) {
  return guarded([&] { CrossEntropyPickBackward(wrap(out), wrap(adj), wrap(in), wrap(pick)); });
}
int mrn_layer_norm(mrn_tensor out, mrn_tensor in, mrn_tensor gamma, const mrn_tensor* beta, float eps) {
  return guarded([&] { LayerNormalization(wrap(out), wrap(in), wrap(gamma), wrapOpt(beta), eps); });
}
int mrn_layer_norm_grad(mrn_tensor gx, mrn_tensor gg, const mrn_tensor* gb, mrn_tensor adj, mrn_tensor y, mrn_tensor x, mrn_tensor gamma, const mrn_tensor* beta, float eps) {
  return guarded([&] {
    LayerNormalizationGrad(wrap(gx), wrap(gg), wrapOpt(gb), wrap(adj), wrap(y), wrap(x), wrap(gamma), wrapOpt(beta), eps);
  });
}
int mrn_residual_layer_norm(mrn_tensor out, mrn_tensor in, mrn_tensor residual, mrn_tensor gamma, mrn_tensor beta, float eps) {
  return guarded([&] { LayerNormalization(wrap(out), wrap(in), wrap(gamma), wrap(beta), eps, wrap(residual)); });
}
int mrn_residual_layer_norm_grad(mrn_tensor gx, mrn_tensor gr, mrn_tensor gg, mrn_tensor gb, mrn_tensor adj, mrn_tensor y, mrn_tensor x, mrn_tensor residual, mrn_tensor gamma, mrn_tensor beta, float eps) {
  return guarded([&] {
    LayerNormalizationGrad(wrap(gx), wrap(gg), wrap(gb), wrap(adj), wrap(y), wrap(x), wrap(gamma), wrap(beta), eps, wrap(residual), wrap(gr));
  });
}
int mrn_multi_head_attention(mrn_tensor out, mrn_tensor probs, mrn_tensor q, mrn_tensor k, mrn_tensor v, const mrn_tensor* mask, int heads, float scale, int exact) {
  return guarded([&] { MultiHeadAttention(wrap(out), wrap(probs), wrap(q), wrap(k), wrap(v), wrapOpt(mask), heads, scale, exact != 0); });
}
int mrn_multi_head_attention_grad(mrn_tensor dq, mrn_tensor dk, mrn_tensor dv, mrn_tensor adj, mrn_tensor out, mrn_tensor probs, mrn_tensor q, mrn_tensor k, mrn_tensor v, int heads, float scale, int exact) {
  return guarded([&] {
    MultiHeadAttentionGrad(wrap(dq), wrap(dk), wrap(dv), wrap(adj), wrap(out), wrap(probs), wrap(q), wrap(k), wrap(v), heads, scale, exact != 0);
  });
}
int mrn_att(mrn_tensor out, mrn_tensor va, mrn_tensor context, mrn_tensor state) {
  return guarded([&] { Att(wrap(out), wrap(va), wrap(context), wrap(state)); });
}
int mrn_att_back(mrn_tensor gva, mrn_tensor gc, mrn_tensor gs, mrn_tensor va, mrn_tensor context, mrn_tensor state, mrn_tensor adj) {
  return guarded([&] { AttBack(wrap(gva), wrap(gc), wrap(gs), wrap(va), wrap(context), wrap(state), wrap(adj)); });
}

int mrn_gru_fast_forward(mrn_tensor out, const mrn_tensor* inputs, int n, int final) {
  return guarded([&] { GRUFastForward(wrap(out), wrapAll(inputs, n), final != 0); });
}
int mrn_gru_fast_backward(const mrn_tensor* outputs, const mrn_tensor* inputs, int n, mrn_tensor adj, int final) {
  return guarded([&] { GRUFastBackward(wrapAll(outputs, 4), wrapAll(inputs, n), wrap(adj), final != 0); });
}
int mrn_lstm_cell_forward(mrn_tensor out, const mrn_tensor* inputs, int n) {
  return guarded([&] { LSTMCellForward(wrap(out), wrapAll(inputs, n)); });
}
int mrn_lstm_output_forward(mrn_tensor out, const mrn_tensor* inputs, int n) {
  return guarded([&] { LSTMOutputForward(wrap(out), wrapAll(inputs, n)); });
}
int mrn_lstm_cell_backward(const mrn_tensor* outputs, const mrn_tensor* inputs, int n, mrn_tensor adj) {
  return guarded([&] { LSTMCellBackward(wrapAll(outputs, 4), wrapAll(inputs, n), wrap(adj)); });
}
int mrn_lstm_output_backward(const mrn_tensor* outputs, const mrn_tensor* inputs, int n, mrn_tensor adj) {
  return guarded([&] { LSTMOutputBackward(wrapAll(outputs, 4), wrapAll(inputs, n), wrap(adj)); });
}
int mrn_highway_forward(mrn_tensor out, mrn_tensor in1, mrn_tensor in2, mrn_tensor t) {
  return guarded([&] { HighwayForward(wrap(out), wrap(in1), wrap(in2), wrap(t)); });
}
int mrn_highway_backward(mrn_tensor o1, mrn_tensor o2, mrn_tensor ot, mrn_tensor in1, mrn_tensor in2, mrn_tensor t, mrn_tensor adj) {
  return guarded([&] { HighwayBackward(wrap(o1), wrap(o2), wrap(ot), wrap(in1), wrap(in2), wrap(t), wrap(adj)); });
}

int mrn_transpose_nd(mrn_tensor out, mrn_tensor in, const int* axes) {
  return guarded([&] { TransposeND(wrap(out), wrap(in), std::vector<int>(axes, axes + in.rank)); });
}
int mrn_concatenate(mrn_tensor out, const mrn_tensor* ins, int n, int axis) {
  return guarded([&] {
    int ax = axis < 0 ? out.rank + axis : axis;
    Concatenate(wrap(out), wrapAll(ins, n), ax);
  });
}
int mrn_deconcatenate(const mrn_tensor* outs, int n, mrn_tensor in, int axis) {
  return guarded([&] {
    int ax = axis < 0 ? in.rank + axis : axis;
    auto o = wrapAll(outs, n);
    Deconcatenate(o, wrap(in), ax);
  });
}
int mrn_copy_rows(mrn_tensor out, mrn_tensor in, const int* idx, size_t n) {
  return guarded([&] { CopyRows(wrap(out), wrap(in), idx, n); });
}
int mrn_paste_rows(mrn_tensor out, mrn_tensor in, const int* idx, size_t n) {
  return guarded([&] { PasteRows(wrap(out), wrap(in), idx, n); });
}
int mrn_shift(mrn_tensor out, mrn_tensor in, const int* shift, int invert) {
  return guarded([&] { Shift(wrap(out), wrap(in), Shape(std::vector<int>(shift, shift + in.rank)), invert != 0); });
}

int mrn_l2norm(mrn_tensor in, float* result) {
  return guarded([&] { *result = L2Norm(wrap(in)); });
}

int mrn_adam_step(mrn_tensor params, mrn_tensor grads, mrn_tensor mt, mrn_tensor vt, float eta, float beta1, float beta2, float eps, int t, float grad_scale, float clip_norm) {
  return guarded([&] {
    AdamArgs a;
    a.eta = eta;
    a.beta1 = beta1;
    a.beta2 = beta2;
    a.eps = eps;
    a.denom1 = (float)(1 - std::pow((double)beta1, (double)t));
    a.denom2 = (float)(1 - std::pow((double)beta2, (double)t));
    a.gradScale = grad_scale;
    a.clipNorm = clip_norm;
    Tensor normSq = nullptr;
    Ptr<TensorAllocator> scratch;
    if(clip_norm > 0) {
      scratch = New<TensorAllocator>(device::getDevice());
      scratch->reserveExact(256);
      scratch->allocate(normSq, Shape{1, 1});
      SumSquares(normSq, wrap(grads));
    }
    AdamUpdate(wrap(params), wrap(grads), wrap(mt), wrap(vt), a, normSq);
    if(scratch)
      device::synchronize();  // scratch is released on return
  });
}

namespace {
// clip-norm scratch of the stand-alone optimizer entry points
struct NormScratch {
  Ptr<TensorAllocator> alloc;
  Tensor normSq;
  NormScratch(Tensor grads, float clipNorm) {
    if(clipNorm > 0) {
      alloc = New<TensorAllocator>(device::getDevice());
      alloc->reserveExact(256);
      alloc->allocate(normSq, Shape{1, 1});
      SumSquares(normSq, grads);
    }
  }
  ~NormScratch() {
    if(alloc)
      device::synchronize();  // scratch is released on return
  }
};
}  // namespace

int mrn_sgd_step(mrn_tensor params, mrn_tensor grads, float eta, float grad_scale, float clip_norm) {
  return guarded([&] {
    NormScratch ns(wrap(grads), clip_norm);
    SgdUpdate(wrap(params), wrap(grads), eta, grad_scale, clip_norm, ns.normSq);
  });
}
int mrn_adagrad_step(mrn_tensor params, mrn_tensor grads, mrn_tensor gt, float eta, float eps, float grad_scale, float clip_norm) {
  return guarded([&] {
    NormScratch ns(wrap(grads), clip_norm);
    AdagradUpdate(wrap(params), wrap(grads), wrap(gt), eta, eps, grad_scale, clip_norm, ns.normSq);
  });
}
int mrn_dropout(mrn_tensor mask, float drop_prob, unsigned long long seed) {
  return guarded([&] { Dropout(wrap(mask), drop_prob, (uint64_t)seed, nullptr); });
}

// ---------------------------------------------------------------------------
// training-step driver
// ---------------------------------------------------------------------------
int mrn_trainer_create(void** trainer, const char* options, int dev, int rank, int nranks) {
  return guarded([&] {
    auto t = new Trainer();
    t->options = defaultOptions();
    Options user(options ? options : "");
    t->options->overwrite(user);
    t->device = dev;
    t->rank = rank;
    t->nranks = nranks;
    Config::seed = t->options->get<size_t>("seed");
    device::setDevice(dev);
    if(t->options->get<std::string>("graph-group", "sync") == "async")
      t->async = New<AsyncGraphGroup>(t->options, dev, rank, nranks);
    else if(nranks > 1)
      t->sync = New<SyncGraphGroup>(t->options, dev, rank, nranks);
    else
      t->single = New<SingletonGraph>(t->options, dev);
    auto gemm = t->worker().graph()->getBackend()->getGemmHandle();
    setGemmMode(gemm, (GemmMode)t->options->get<int>("gemm-mode", 0));
    auto voc = t->options->get<std::vector<int>>("dim-vocabs");
    t->corpus = New<data::SyntheticCorpus>(voc[0], voc[1], (uint32_t)t->options->get<int>("data-seed", 1111));
    *trainer = t;
  });
}
int mrn_trainer_destroy(void* trainer) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    device::setDevice(t->device);
    device::synchronize();
    delete t;
  });
}

int mrn_trainer_set_batch(void* trainer, int B, int Ts, const int64_t* srcIdx, const float* srcMask, int Tt, const int64_t* trgIdx, const float* trgMask) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    auto mk = [&](int T, const int64_t* idx, const float* mask) {
      auto sb = New<data::SubBatch>(B, T);
      size_t words = 0;
      for(size_t i = 0; i < (size_t)B * T; ++i) {
        sb->indices()[i] = (Word)idx[i];
        sb->mask()[i] = mask[i];
        if(mask[i] != 0)
          words++;
      }
      sb->setWords(words);
      return sb;
    };
    t->batch = New<data::CorpusBatch>(std::vector<Ptr<data::SubBatch>>{mk(Ts, srcIdx, srcMask), mk(Tt, trgIdx, trgMask)});
  });
}

int mrn_trainer_next_synthetic_batch(void* trainer, int B, int Ls, int Lt, int padded, int splitRank, int splitN) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    auto full = t->corpus->next(B, Ls, Lt, padded != 0);
    if(splitN > 1)
      t->batch = full->split(splitN)[splitRank];
    else
      t->batch = full;
  });
}

// ---- text corpus in front of the hot path (data/corpus.h; reference src/data/{vocab,corpus}.cpp, batch_generator.h) ----
int mrn_trainer_open_corpus(void* trainer, const char* srcPath, const char* trgPath, const char* vocabSrc, const char* vocabTrg, const char* options) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    auto o = t->options->clone();
    o->overwrite(Options(options ? options : ""));
    std::vector<std::string> paths{srcPath, trgPath};
    std::vector<std::string> vocabPaths{vocabSrc ? vocabSrc : "", vocabTrg ? vocabTrg : ""};
    auto dims = t->options->get<std::vector<int>>("dim-vocabs");
    std::vector<Ptr<data::Vocab>> vocabs;
    for(size_t i = 0; i < 2; ++i) {
      auto v = New<data::Vocab>();
      int size = v->loadOrCreate(vocabPaths[i], paths[i], dims[i]);  // ids >= dim-vocabs[i] are dropped (-> <unk>), as in the reference
      ABORT_IF(size > dims[i], "vocabulary", i, "has", size, "entries, the model was created with dim-vocabs", dims[i]);
      vocabs.push_back(v);
    }
    t->textBatches = New<data::PrefetchingBatchGenerator>(New<data::Corpus>(paths, vocabs, o), o);
  });
}
// *hasBatch = 0 at the end of an epoch (the generator restarts: the next call begins the next epoch)
int mrn_trainer_next_corpus_batch(void* trainer, int* hasBatch) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->textBatches, "mrn_trainer_next_corpus_batch: call mrn_trainer_open_corpus first");
    auto b = t->textBatches->next();
    if(b) {
      t->batch = t->nranks > 1 ? b->split(t->nranks)[t->rank] : b;
      *hasBatch = 1;
    } else {
      t->textBatches->restart();
      *hasBatch = 0;
    }
  });
}
// cross-entropy validation on a held-out corpus (training/validator.h; reference training/validator.h:108-176)
int mrn_trainer_validate(void* trainer, const char* srcPath, const char* trgPath, const char* vocabSrc, const char* vocabTrg, const char* options, float* metric, float* costSum,
                         size_t* sentences, size_t* targetWords) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    auto graph = t->worker().graph();
    ABORT_IF(graph->params()->size() == 0, "mrn_trainer_validate: the model has no parameters yet (run a step or load a checkpoint first)");
    auto o = t->options->clone();
    o->overwrite(Options(options ? options : ""));
    std::vector<std::string> paths{srcPath, trgPath};
    std::vector<std::string> vocabPaths{vocabSrc ? vocabSrc : "", vocabTrg ? vocabTrg : ""};
    auto dims = t->options->get<std::vector<int>>("dim-vocabs");
    std::vector<Ptr<data::Vocab>> vocabs;
    for(size_t i = 0; i < 2; ++i) {
      auto v = New<data::Vocab>();
      v->loadOrCreate(vocabPaths[i], paths[i], dims[i]);
      vocabs.push_back(v);
    }
    CrossEntropyValidator validator(vocabs, o);
    float sum = 0;
    size_t n = 0, w = 0;
    *metric = validator.validate(graph, paths, &sum, &n, &w);
    if(costSum)
      *costSum = sum;
    if(sentences)
      *sentences = n;
    if(targetWords)
      *targetWords = w;
  });
}
int mrn_nth_element_ranges(mrn_tensor scores, const int* rangeFirst, const int* cumN, int ranges, float* outCosts, unsigned* outKeys) {
  return guarded([&] {
    std::vector<int> first(rangeFirst, rangeFirst + ranges + 1), cum(cumN, cumN + ranges + 1);
    std::vector<float> costs;
    std::vector<unsigned> keys;
    NthElementRanges(wrap(scores), first, cum, costs, keys);
    std::copy(costs.begin(), costs.end(), outCosts);
    std::copy(keys.begin(), keys.end(), outKeys);
  });
}
int mrn_nth_element_logsoftmax(mrn_tensor logits, const float* prevCosts, int dimBatch, int beam, int n, int first, int suppressWord, float* outCosts, unsigned* outKeys) {
  return guarded([&] {
    std::vector<float> prev(prevCosts, prevCosts + (size_t)(first ? 1 : beam) * dimBatch);
    std::vector<float> costs;
    std::vector<unsigned> keys;
    NthElementLogSoftmax(wrap(logits), prev, dimBatch, beam, n, first != 0, suppressWord, costs, keys);
    std::copy(costs.begin(), costs.end(), outCosts);
    std::copy(keys.begin(), keys.end(), outKeys);
  });
}
// beam search over the source side of the current batch (translator/beam_search.h; reference translator/beam_search.h:91-225)
int mrn_trainer_translate(void* trainer, const char* options, int nBest, int maxLen, int64_t* words, int* lengths, float* scores, float* rawScores) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    auto graph = t->worker().graph();
    ABORT_IF(!t->batch, "mrn_trainer_translate: no current batch");
    ABORT_IF(graph->params()->size() == 0, "mrn_trainer_translate: the model has no parameters yet (run a step or load a checkpoint first)");
    ABORT_IF(nBest < 1 || maxLen < 1, "mrn_trainer_translate: n_best and max_len must be positive");
    Translator translator(t->options, New<Options>(options ? options : ""));
    auto results = translator.translate(graph, t->batch, (size_t)nBest);
    for(size_t s = 0; s < results.size(); ++s)
      for(int r = 0; r < nBest; ++r) {
        size_t slot = s * nBest + r;
        bool have = r < (int)results[s].size();
        int len = have ? (int)results[s][r].words.size() : -1;
        lengths[slot] = len;
        scores[slot] = have ? results[s][r].cost : 0.f;
        if(rawScores)
          rawScores[slot] = have ? results[s][r].rawCost : 0.f;
        for(int k = 0; k < maxLen; ++k)
          words[slot * maxLen + k] = (have && k < len) ? (int64_t)results[s][r].words[k] : -1;
      }
  });
}
// text file -> translations in corpus order (translator/translator.h; reference translator/translator.h:22-108)
int mrn_trainer_translate_file(void* trainer, const char* srcPath, const char* vocabSrc, const char* vocabTrg, const char* options, const char* outPath, size_t* sentences) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    auto graph = t->worker().graph();
    ABORT_IF(graph->params()->size() == 0, "mrn_trainer_translate_file: the model has no parameters yet (run a step or load a checkpoint first)");
    ABORT_IF(!vocabSrc || !vocabTrg, "mrn_trainer_translate_file: both vocabularies are required");
    auto dims = t->options->get<std::vector<int>>("dim-vocabs");
    auto vs = New<data::Vocab>(), vt = New<data::Vocab>();
    vs->loadOrCreate(vocabSrc, srcPath, dims[0]);
    vt->load(vocabTrg, dims[1]);
    Translator translator(t->options, New<Options>(options ? options : ""));
    size_t n = translator.translateFile(graph, srcPath, vs, vt, outPath);
    if(sentences)
      *sentences = n;
  });
}
// current batch as host arrays (time-major [T, B], the SubBatch layout); side 0 = source, 1 = target
int mrn_trainer_get_batch(void* trainer, int side, int64_t* indices, float* mask, size_t capacity, int* batchSize, int* width) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->batch || side < 0 || side >= (int)t->batch->sets(), "mrn_trainer_get_batch: no such sub-batch");
    auto sb = (*t->batch)[side];
    *batchSize = (int)sb->batchSize();
    *width = (int)sb->batchWidth();
    size_t n = sb->batchSize() * sb->batchWidth();
    if(indices && mask) {
      ABORT_IF(capacity < n, "mrn_trainer_get_batch: buffer too small");
      for(size_t i = 0; i < n; ++i) {
        indices[i] = (int64_t)sb->indices()[i];
        mask[i] = sb->mask()[i];
      }
    }
  });
}

int mrn_trainer_compute_gradients(void* trainer, int keepLogits) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->batch, "no batch set");
    if(t->sync) {
      ABORT_IF(keepLogits, "keep_logits is only supported on the single-process trainer");
      t->sync->computeGradients(t->batch);
    } else {
      t->worker().computeGradients(t->batch, keepLogits != 0);
    }
    if(t->worker().lastStepReplayed())
      t->replays++;
  });
}
int mrn_trainer_update(void* trainer) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->single, "mrn_trainer_update: sharded trainers use mrn_trainer_update_shard");
    t->single->optimizer()->update(t->worker().graph());
  });
}
int mrn_trainer_update_shard(void* trainer) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->sync, "mrn_trainer_update_shard needs nranks > 1");
    t->sync->updateShard();
  });
}
// ---- peer-memory exchange: CUDA IPC handles of {parameter arena, gradient arena, signal pad} ----
int mrn_trainer_ipc_export(void* trainer, unsigned char* handles, size_t capacity) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->sync, "mrn_trainer_ipc_export needs nranks > 1");
    ABORT_IF(capacity < 3 * device::ipcHandleBytes(), "mrn_trainer_ipc_export: buffer too small");
    device::setDevice(t->device);
    void* ptrs[3] = {t->sync->flatParams()->data(), t->sync->flatGrads()->data(), t->sync->signalPad()};
    for(int i = 0; i < 3; ++i)
      device::ipcExport(ptrs[i], handles + i * device::ipcHandleBytes());
  });
}
int mrn_trainer_ipc_import(void* trainer, const unsigned char* allHandles, int nranks) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->sync || nranks != t->nranks || nranks > 8, "mrn_trainer_ipc_import: rank count mismatch (max 8 ranks per node)");
    device::setDevice(t->device);
    PeerTable tab[3] = {};
    void* own[3] = {t->sync->flatParams()->data(), t->sync->flatGrads()->data(), t->sync->signalPad()};
    size_t hb = device::ipcHandleBytes();
    for(int r = 0; r < nranks; ++r)
      for(int i = 0; i < 3; ++i)
        tab[i].ptr[r] = r == t->rank ? own[i] : device::ipcOpen(allHandles + ((size_t)r * 3 + i) * hb);
    t->sync->setPeers(tab[0], tab[1], tab[2]);
  });
}
int mrn_trainer_update_peer(void* trainer) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->sync, "mrn_trainer_update_peer needs nranks > 1");
    t->sync->exchangeUpdatePeer();
  });
}

// ---- asynchronous parameter server (AsyncGraphGroup; options "graph-group=async") ----
int mrn_trainer_async_init(void* trainer) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->async || !t->batch, "mrn_trainer_async_init needs graph-group=async and a current batch");
    t->async->init(t->batch);
  });
}
int mrn_trainer_async_export(void* trainer, unsigned char* handle, size_t capacity) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->async || !t->async->masterBlock(), "mrn_trainer_async_export: call mrn_trainer_async_init first");
    ABORT_IF(capacity < device::ipcHandleBytes(), "mrn_trainer_async_export: buffer too small");
    device::setDevice(t->device);
    device::ipcExport(t->async->masterBlock(), handle);
  });
}
int mrn_trainer_async_import(void* trainer, const unsigned char* allHandles, int nranks) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->async || nranks != t->nranks || nranks > 8, "mrn_trainer_async_import: rank count mismatch (max 8 ranks per node)");
    device::setDevice(t->device);
    PeerTable tab = {};
    for(int r = 0; r < nranks; ++r)
      tab.ptr[r] = r == t->rank ? t->async->masterBlock() : device::ipcOpen(allHandles + (size_t)r * device::ipcHandleBytes());
    t->async->setPeers(tab);
  });
}
int mrn_trainer_async_update(void* trainer) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->async || !t->batch, "mrn_trainer_async_update needs graph-group=async and a current batch");
    t->async->update(t->batch);
  });
}
int mrn_trainer_async_fetch(void* trainer) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->async, "mrn_trainer_async_fetch needs graph-group=async");
    t->async->fetchParams();
  });
}

// ---- checkpoint / resume (training/checkpoint.h; reference expression_graph.h:442-502, encdec.h:201-229) ----
int mrn_trainer_save(void* trainer, const char* path, int withOptimizer) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    device::setDevice(t->device);
    device::synchronize();
    auto graph = t->worker().graph();
    ABORT_IF(graph->params()->size() == 0, "mrn_trainer_save: the model has no parameters yet (run a step or load a checkpoint first)");
    checkpoint::saveModel(graph, t->options, path);
    if(withOptimizer) {
      ABORT_IF(!t->single, "mrn_trainer_save: optimizer state is saved by the single-process trainer (shard optimizers keep per-rank state)");
      auto adam = std::dynamic_pointer_cast<Adam>(t->single->optimizer());
      ABORT_IF(!adam, "mrn_trainer_save: optimizer state is implemented for adam");
      checkpoint::saveAdam(graph, adam, std::string(path) + ".optimizer.npz");
    }
  });
}
int mrn_trainer_load(void* trainer, const char* path, int withOptimizer) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    device::setDevice(t->device);
    auto graph = t->worker().graph();
    ABORT_IF(graph->params()->size() != 0, "mrn_trainer_load: load a checkpoint into a fresh trainer, before its first step");
    Options stored = checkpoint::loadModel(graph, path);
    // the stored model description must describe the model this trainer builds
    for(auto& key : checkpoint::modelFeatures())
      if(stored.has(key) && t->options->has(key))
        ABORT_IF(stored.get<std::string>(key) != t->options->get<std::string>(key), "mrn_trainer_load: checkpoint was trained with a different", key, ":",
                 stored.get<std::string>(key), "vs", t->options->get<std::string>(key));
    if(withOptimizer) {
      ABORT_IF(!t->single, "mrn_trainer_load: optimizer state is restored by the single-process trainer");
      auto adam = std::dynamic_pointer_cast<Adam>(t->single->optimizer());
      ABORT_IF(!adam, "mrn_trainer_load: optimizer state is implemented for adam");
      checkpoint::loadAdam(graph, adam, std::string(path) + ".optimizer.npz");
    }
  });
}

int mrn_trainer_cost(void* trainer, float* cost) {
  return guarded([&] { *cost = ((Trainer*)trainer)->worker().cost(); });
}

int mrn_trainer_params(void* trainer, float** ptr, size_t* elements) {
  return guarded([&] {
    auto p = ((Trainer*)trainer)->worker().graph()->params()->vals();
    *ptr = p->data();
    *elements = p->size();
  });
}
int mrn_trainer_grads(void* trainer, float** ptr, size_t* elements) {
  return guarded([&] {
    auto g = ((Trainer*)trainer)->worker().graph()->params();
    g->allocateBackward();
    auto p = g->grads();
    *ptr = p->data();
    *elements = p->size();
  });
}
int mrn_trainer_shard_grads(void* trainer, float** ptr, size_t* elements) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->sync, "mrn_trainer_shard_grads needs nranks > 1");
    auto p = t->sync->shardGrads();
    *ptr = p->data();
    *elements = p->size();
  });
}

int mrn_trainer_get_tensor(void* trainer, const char* name, int wantGrad, float* host, size_t capacity, size_t* elements) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    std::string n = name;
    Tensor src;
    if(n == "logits") {
      auto l = t->worker().logits();
      ABORT_IF(!l, "no logits kept: call mrn_trainer_compute_gradients(keep_logits=1)");
      src = wantGrad ? l->grad() : l->val();
    } else {
      auto p = t->worker().graph()->params()->get(n);
      ABORT_IF(!p, "unknown parameter", n);
      src = wantGrad ? p->grad() : p->val();
    }
    ABORT_IF(!src, "tensor not allocated", n);
    *elements = src->size();
    if(host) {
      ABORT_IF(capacity < src->size(), "buffer too small for", n);
      std::vector<float> v;
      src->get(v);
      std::memcpy(host, v.data(), v.size() * sizeof(float));
    }
  });
}

int mrn_trainer_param_names(void* trainer, char* buffer, size_t capacity, size_t* needed) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    std::string s;
    for(auto p : *t->worker().graph()->params()) {
      s += p->name() + " " + std::to_string(p->shape().size());
      for(auto d : p->shape())
        s += " " + std::to_string(d);
      s += "\n";
    }
    *needed = s.size() + 1;
    if(buffer && capacity >= s.size() + 1)
      std::memcpy(buffer, s.c_str(), s.size() + 1);
  });
}

int mrn_trainer_batch_words(void* trainer, size_t* srcWords, size_t* totalWords) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    ABORT_IF(!t->batch, "no batch set");
    *srcWords = t->batch->words();
    *totalWords = t->batch->wordsTotal();
  });
}

int mrn_trainer_stats(void* trainer, size_t* tapeNodes, size_t* plans, size_t* replays, size_t* workspaceBytes) {
  return guarded([&] {
    auto t = (Trainer*)trainer;
    *tapeNodes = t->worker().graph()->numNodes();
    *plans = t->worker().replay().size();
    *replays = t->replays;
    *workspaceBytes = t->worker().graph()->allocator()->peak();
  });
}

int mrn_trainer_graph_kernels(void* trainer, size_t* kernels) {
  return guarded([&] { *kernels = ((Trainer*)trainer)->worker().replay().lastPlanKernels(); });
}

}  // extern "C"
