// Optimizers over ONE flat parameter tensor (whole model or a shard of it).
//
// Semantics follow the reference (src/optimizers/optimizers.h:14-162,
// optimizers.cu:7-107, clippers.cu:12-17):
//   update(params, grads, mult):  clip grads by their L2 norm (if >= c: g *= c/|g|),
//   then Sgd / Adagrad / Adam with
//     m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
//     p -= eta*mult * (m / (1-b1^t)) / (sqrt(v / (1-b2^t)) + eps).
// The reference runs norm (with a host round trip), scale, and 1-3 Element
// passes + a stream sync.  Here: one asynchronous sum-of-squares reduction
// into a device scalar, then ONE fused kernel that applies clip factor,
// optional 1/N gradient averaging and the update (kernels/tensor_operators.h:
// AdamUpdate & co).  Nothing synchronises.
#pragma once

#include <cmath>
#include <functional>

#include "common/options.h"
#include "graph/expression_graph.h"
#include "kernels/tensor_operators.h"

namespace marian {

class OptimizerBase {
public:
  OptimizerBase(float eta, float clipNorm) : eta_(eta), clipNorm_(clipNorm) {}
  virtual ~OptimizerBase() {}

  void update(Ptr<ExpressionGraph> graph, float multiplyFactor = 1.0f) {
    Tensor p = graph->params()->vals();
    Tensor g = graph->params()->grads();
    // BF16S GEMM mode: the update kernel also refreshes the bf16 copy of the parameter arena the
    // tensor-core products read (kernels/shadow.h); optimizers that cannot leave it stale and the
    // next step converts the arena in one pass (gemmPrepareStep)
    auto gemm = graph->getBackend()->getGemmHandle();
    shadowOut_ = gemmParamShadowFor(gemm, p);
    shadowWritten_ = false;
    update(p, g, multiplyFactor);
    gemmParamsUpdated(gemm, shadowWritten_);
    shadowOut_ = nullptr;
  }

  // gradScale multiplies every gradient element first (1/N averaging of a
  // summed shard); clipping then acts on the scaled gradient, as the reference
  // clips the already averaged shard (graph_group_sync.cu:130-142).
  void update(Tensor params, Tensor grads, float multiplyFactor = 1.0f, float gradScale = 1.0f) {
    multiplyFactor_ = multiplyFactor;
    Tensor normSq = nullptr;
    if(clipNorm_ > 0) {
      ensureScratch(params->getDevice());
      SumSquares(normSq_, grads);
      normSq = normSq_;
    }
    updateImpl(params, grads, gradScale, normSq, nullptr);
  }

  // Peer-memory exchange variant (training/graph_group.h, kernels/exchange.cu): `grads` is the
  // summed gradient shard and normSqScratch() already holds its sum of squares (both produced by
  // PeerGatherReduce); the update kernel also stores the new parameters into the peers' arenas.
  void updateShardWithPeers(Tensor params, Tensor grads, float gradScale, const PeerStores& peers) {
    multiplyFactor_ = 1.f;
    ensureScratch(params->getDevice());
    updateImpl(params, grads, gradScale, clipNorm_ > 0 ? normSq_ : nullptr, &peers);
  }
  Tensor normSqScratch(int device) {
    ensureScratch(device);
    return normSq_;
  }

  void setLearnRate(float eta) { eta_ = eta; }
  float learnRate() const { return eta_; }

  // device scalar holding sum(g^2) of the last update (unscaled); for logging/tests
  Tensor lastNormSq() { return normSq_; }

protected:
  virtual void updateImpl(Tensor params, Tensor grads, float gradScale, Tensor normSq, const PeerStores* peers) = 0;

  void ensureScratch(int device) {
    if(!normSq_) {
      scratch_ = New<TensorAllocator>(device);
      scratch_->reserveExact(256);
      scratch_->allocate(normSq_, Shape{1, 1});
    }
  }

  float eta_;
  float clipNorm_;
  float multiplyFactor_{1.f};
  void* shadowOut_{nullptr};   // bf16 destination for the updated parameters (whole-arena updates only)
  bool shadowWritten_{false};  // set by an updateImpl that honoured shadowOut_
  Ptr<TensorAllocator> scratch_;
  Tensor normSq_;
};

class Sgd : public OptimizerBase {
public:
  Sgd(float eta, float clipNorm) : OptimizerBase(eta, clipNorm) {}

private:
  void updateImpl(Tensor params, Tensor grads, float gradScale, Tensor normSq, const PeerStores* peers) {
    ABORT_IF(peers, "the peer-memory exchange is fused with Adam only; use the collective exchange with sgd");
    SgdUpdate(params, grads, multiplyFactor_ * eta_, gradScale, clipNorm_, normSq);
  }
};

class Adagrad : public OptimizerBase {
public:
  Adagrad(float eta, float clipNorm, float eps = 1e-8f) : OptimizerBase(eta, clipNorm), eps_(eps) {}

private:
  void updateImpl(Tensor params, Tensor grads, float gradScale, Tensor normSq, const PeerStores* peers) {
    ABORT_IF(peers, "the peer-memory exchange is fused with Adam only; use the collective exchange with adagrad");
    if(!gt_) {
      alloc_ = New<TensorAllocator>(params->getDevice());
      alloc_->reserveExact(params->memory()->size());
      alloc_->allocate(gt_, Shape{1, (int)params->size()});
      gt_->set(0);
    }
    AdagradUpdate(params, grads, gt_, multiplyFactor_ * eta_, eps_, gradScale, clipNorm_, normSq);
  }
  float eps_;
  Ptr<TensorAllocator> alloc_;
  Tensor gt_;
};

class Adam : public OptimizerBase {
public:
  Adam(float eta, float clipNorm, float beta1 = 0.9f, float beta2 = 0.999f, float eps = 1e-8f)
      : OptimizerBase(eta, clipNorm), beta1_(beta1), beta2_(beta2), eps_(eps) {}

  Tensor mt() { return mt_; }
  Tensor vt() { return vt_; }
  size_t steps() const { return t_; }
  // Piece-wise shard updates (SyncGraphGroup's overlapped exchange): ONE call per update - allocates the moments for
  // `stateElements` parameters, counts the step and returns the arguments for AdamUpdatePieces of all phases.
  AdamArgs beginPieceStep(size_t stateElements, int device, float gradScale) {
    if(!mt_) {
      alloc_ = New<TensorAllocator>(device);
      alloc_->reserveExact(2 * alloc_->capacity(Shape{1, (int)stateElements}));
      alloc_->allocate(mt_, Shape{1, (int)stateElements});
      mt_->set(0);
      alloc_->allocate(vt_, Shape{1, (int)stateElements});
      vt_->set(0);
    }
    ABORT_IF(mt_->size() != stateElements, "Adam state has a different size than the shard");
    t_++;
    AdamArgs a = hyper(gradScale);
    a.denom1 = (float)(1 - std::pow((double)beta1_, (double)t_));
    a.denom2 = (float)(1 - std::pow((double)beta2_, (double)t_));
    return a;
  }
  // resume (training/checkpoint.h): `fill` writes the saved moments once the flat state exists
  void restoreOnAllocation(size_t steps, std::function<void(Tensor, Tensor)> fill) {
    pendingSteps_ = steps;
    pendingFill_ = fill;
    if(mt_) {
      pendingFill_(mt_, vt_);
      t_ = pendingSteps_;
      pendingFill_ = nullptr;
    }
  }
  // hyper-parameters for update kernels that keep their own state (asynchronous parameter server)
  AdamArgs hyper(float gradScale = 1.f) const {
    AdamArgs a;
    a.eta = eta_;
    a.beta1 = beta1_;
    a.beta2 = beta2_;
    a.eps = eps_;
    a.denom1 = a.denom2 = 1.f;
    a.gradScale = gradScale;
    a.clipNorm = clipNorm_;
    return a;
  }

private:
  void updateImpl(Tensor params, Tensor grads, float gradScale, Tensor normSq, const PeerStores* peers) {
    if(!mt_) {
      alloc_ = New<TensorAllocator>(params->getDevice());
      alloc_->reserveExact(2 * alloc_->capacity(Shape{1, (int)params->size()}));
      alloc_->allocate(mt_, Shape{1, (int)params->size()});
      mt_->set(0);
      alloc_->allocate(vt_, Shape{1, (int)params->size()});
      vt_->set(0);
      if(pendingFill_) {
        pendingFill_(mt_, vt_);
        t_ = pendingSteps_;
        pendingFill_ = nullptr;
      }
    }
    t_++;
    AdamArgs a;
    a.eta = multiplyFactor_ * eta_;
    a.beta1 = beta1_;
    a.beta2 = beta2_;
    a.eps = eps_;
    a.denom1 = (float)(1 - std::pow((double)beta1_, (double)t_));  // double pow, as std::pow(float, size_t) promotes
    a.denom2 = (float)(1 - std::pow((double)beta2_, (double)t_));
    a.gradScale = gradScale;
    a.clipNorm = clipNorm_;
    a.shadow = shadowOut_;
    shadowWritten_ = shadowOut_ != nullptr;
    AdamUpdate(params, grads, mt_, vt_, a, normSq, peers);
  }

  float beta1_, beta2_, eps_;
  size_t t_{0};
  Ptr<TensorAllocator> alloc_;
  Tensor mt_;
  Tensor vt_;
  size_t pendingSteps_{0};
  std::function<void(Tensor, Tensor)> pendingFill_;
};

// reference: optimizers.cu:85-107
inline Ptr<OptimizerBase> Optimizer(Ptr<Options> options) {
  float lrate = (float)options->get<double>("learn-rate");
  float clipNorm = (float)options->get<double>("clip-norm");
  auto opt = options->get<std::string>("optimizer");
  if(opt == "sgd")
    return New<Sgd>(lrate, clipNorm);
  if(opt == "adagrad")
    return New<Adagrad>(lrate, clipNorm);
  if(opt == "adam")
    return New<Adam>(lrate, clipNorm);
  ABORT("Unknown optimizer:", opt);
}

}  // namespace marian
