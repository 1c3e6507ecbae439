// Multi-input graph nodes.  Forward/backward formulas follow the reference's
// src/graph/node_operators_binary.h (cited per class).
#pragma once

#include "graph/backend.h"
#include "graph/expression_graph.h"
#include "graph/node.h"
#include "graph/node_operators_unary.h"
#include "kernels/tensor_operators.h"

namespace marian {

namespace detail {
inline Shape dotShape(Expr a, Expr b, bool transA, bool transB) {
  auto shapeA = a->shape();
  if(transA) {
    shapeA.set(-2, a->shape()[-1]);
    shapeA.set(-1, a->shape()[-2]);
  }
  auto shapeB = b->shape();
  if(transB) {
    shapeB.set(-2, b->shape()[-1]);
    shapeB.set(-1, b->shape()[-2]);
  }
  Shape outShape = shapeA;
  outShape.set(-1, shapeB[-1]);
  ABORT_IF(shapeA[-1] != shapeB[-2], "matrix product requires dimensions to match", shapeA.toString(), shapeB.toString());
  return outShape;
}
}  // namespace detail

// C = scalar * op(A) op(B); all leading dims of A flattened into rows.
// reference: node_operators_binary.h:13-157.  The four backward transpose cases
// are generated from one table instead of four copies.
class DotNodeOp : public NaryNodeOp {
protected:
  bool transA_;
  bool transB_;
  float scalar_;

  virtual void prod(Tensor C, Tensor A, Tensor B, bool tA, bool tB, float beta) {
    Prod(getBackend()->getGemmHandle(), C, A, B, tA, tB, beta, scalar_);
  }
  // gradient of child i: weight gradients leave the critical path
  void prodGrad(int i, Tensor A, Tensor B, bool tA, bool tB) {
    offCriticalPath(child(i), [&] { prod(child(i)->grad(), A, B, tA, tB, 1.f); });
  }

public:
  DotNodeOp(Expr a, Expr b, bool transA, bool transB, float scalar)
      : NaryNodeOp({a, b}, detail::dotShape(a, b, transA, transB)), transA_(transA), transB_(transB), scalar_(scalar) {
    a->requestValShadow();  // BF16S GEMM mode: operands and this node's adjoint as bf16 copies
    b->requestValShadow();
    wantAdjShadow_ = true;
  }

  NodeOps forwardOps() { return {NodeOp(prod(val_, child(0)->val(), child(1)->val(), transA_, transB_, 0.f))}; }
  bool readsChildViaProduct(size_t) const { return true; }

  NodeOps backwardOps() {
    // D = adj, A = child0, B = child1
    if(!transA_ && transB_)  // C = A B^T : dA += D B ; dB += D^T A
      return {NodeOp(prodGrad(0, adj_, child(1)->val(), false, false)), NodeOp(prodGrad(1, adj_, child(0)->val(), true, false))};
    if(transA_ && !transB_)  // C = A^T B : dA += B D^T ; dB += A D
      return {NodeOp(prodGrad(0, child(1)->val(), adj_, false, true)), NodeOp(prodGrad(1, child(0)->val(), adj_, false, false))};
    if(transA_ && transB_)  // C = A^T B^T : dA += B^T D^T ; dB += D^T A^T
      return {NodeOp(prodGrad(0, child(1)->val(), adj_, true, true)), NodeOp(prodGrad(1, adj_, child(0)->val(), true, true))};
    // C = A B : dA += D B^T ; dB += A^T D
    return {NodeOp(prodGrad(0, adj_, child(1)->val(), false, true)), NodeOp(prodGrad(1, child(0)->val(), adj_, true, false))};
  }

  virtual size_t hash() {
    if(!hash_) {
      hash_ = NaryNodeOp::hash();
      hash_combine(hash_, transA_);
      hash_combine(hash_, transB_);
      hash_combine(hash_, scalar_);
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<DotNodeOp>(node);
    return cnode && transA_ == cnode->transA_ && transB_ == cnode->transB_ && scalar_ == cnode->scalar_;
  }
  const std::string type() { return "dot"; }
};

// Batched over the leading dims (attention scores / contexts).
// reference: node_operators_binary.h:221-369
class DotBatchedNodeOp : public DotNodeOp {
protected:
  virtual void prod(Tensor C, Tensor A, Tensor B, bool tA, bool tB, float beta) {
    ProdBatched(getBackend()->getGemmHandle(), C, A, B, tA, tB, beta, scalar_);
  }

public:
  DotBatchedNodeOp(Expr a, Expr b, bool transA, bool transB, float scalar) : DotNodeOp(a, b, transA, transB, scalar) {}
  const std::string type() { return "bdot"; }
};

// x W + b.  reference: :159-218 (Prod, then Add(_1, val, bias) as a 2nd kernel;
// here the bias is applied in the GEMM epilogue).  Backward: dx += D W^T,
// dW += x^T D, db += column sums of D.
struct AffineNodeOp : public NaryNodeOp {
  AffineNodeOp(const std::vector<Expr>& nodes) : NaryNodeOp(nodes, newShape(nodes)) {
    nodes[0]->requestValShadow();  // BF16S GEMM mode: operands and this node's adjoint as bf16 copies
    nodes[1]->requestValShadow();
    wantAdjShadow_ = true;
  }

  static Shape newShape(const std::vector<Expr>& nodes) {
    Shape shape1 = nodes[0]->shape();
    Shape shape2 = nodes[1]->shape();
    ABORT_IF(shape1[-1] != shape2[-2], "matrix product requires dimensions to match");
    shape1.set(-1, shape2[-1]);
    return shape1;
  }

  NodeOps forwardOps() {
    return {NodeOp(if(!forwardDone_) ProdAffine(getBackend()->getGemmHandle(), val_, child(0)->val(), child(1)->val(), child(2)->val()))};
  }
  // Projections of the SAME input that follow each other on the tape (key / value / query of an attention block): their
  // values come out of ONE product launch (kernels/gemm.cu ProdSharedA); the partners are allocated here and marked done.
  bool forwardDone_{false};
  bool weightGradDone_{false};
  static std::vector<AffineNodeOp*> sameInputGroup(AffineNodeOp* first, const std::vector<Expr>& upcoming) {
    std::vector<AffineNodeOp*> group{first};
    for(auto& u : upcoming) {
      auto* a = dynamic_cast<AffineNodeOp*>(u.get());
      if(!a || a->child(0) != first->child(0) || a->shape() != first->shape() || a->child(1)->shape() != first->child(1)->shape() || a->child(2)->shape() != first->child(2)->shape())
        break;
      group.push_back(a);
    }
    return group;
  }
  void fuseForward(const std::vector<Expr>& upcoming) {
    if(forwardDone_)
      return;
    auto group = sameInputGroup(this, upcoming);
    if(group.size() < 2)
      return;
    for(auto* a : group)
      if(a->forwardDone_ || a->concurrent() || !a->child(1)->val() || !a->child(2)->val())
        return;
    std::vector<Tensor> vals, weights, biases;
    for(auto* a : group) {
      a->allocate();
      vals.push_back(a->val_);
      weights.push_back(a->child(1)->val());
      biases.push_back(a->child(2)->val());
    }
    if(ProdSharedA(getBackend()->getGemmHandle(), vals, child(0)->val(), weights, biases, false, 0.f))
      for(auto* a : group)
        a->forwardDone_ = true;
  }
  // Projections of the SAME input that follow each other in the backward sweep (query / key / value
  // of an attention block): dX = sum_g adj_g W_g^T is issued as one K-grouped product instead of a
  // chain of dependent accumulating products (kernels/gemm.cu ProdGroupedNT); the partners are
  // marked so that their own input-gradient closure does nothing.
  bool inputGradDone_{false};
  bool biasGradDone_{false};  // the input-gradient product also delivered the column sums of adj (= the bias gradient)
  bool fuseBias() { return child(2)->trainable() && ProdColumnSumsFusable(getBackend()->getGemmHandle(), adj_); }
  bool readsChildViaProduct(size_t i) const { return i < 2; }
  // BF16S GEMM mode: the adjoint of a projection with ONE consumer is written once and read only by this node's own
  // backward products (input gradient, weight gradient, bias gradient from the streamed bf16 tiles) - its writer may
  // then leave just the bf16 copy (kernels/shadow.h: shadowOnly).  Needs every one of those products on the direct
  // bf16 path (16-byte rows) and both gradients actually taken from the products.
  void set_zero_adjoint() {
    const bool fresh = !adj_;
    Node::set_zero_adjoint();
    if(fresh && adj_ && consumers_ == 1 && getGemmMode(getBackend()->getGemmHandle()) == GemmMode::BF16S && child(0)->trainable() && child(1)->trainable()
       && fuseBias() && (child(0)->shape()[-1] % 8) == 0 && (shape_[-1] % 8) == 0)
      adj_->memory()->shadowOnly = true;
  }
  void fuseBackward(const std::vector<Expr>& upcoming) {
    if(inputGradDone_ || !adj_ || !child(0)->trainable())
      return;
    // affine AFTER swish (second layer of the feed-forward block), the swish node runs next and nothing
    // else has contributed to its adjoint: dH (+)= (adj W^T) o swish'(H) comes out of the product's
    // epilogue; the swish node's adjoint and its element-wise backward kernel are skipped.
    if(!upcoming.empty() && upcoming[0] == child(0)) {
      auto* sw = dynamic_cast<SwishNodeOp*>(child(0).get());
      if(sw && !sw->backwardDone_ && sw->child(0)->trainable() && sw->grad() && sw->grad()->isLazyZero()
         && ProdSwishGradFusable(getBackend()->getGemmHandle(), sw->child(0)->val(), adj_, child(1)->val(), sw->child(0)->val())) {
        sw->child(0)->set_zero_adjoint();
        biasGradDone_ = fuseBias();
        ProdSwishGradNT(getBackend()->getGemmHandle(), sw->child(0)->grad(), adj_, child(1)->val(), sw->child(0)->val(), 1.0, biasGradDone_ ? child(2)->grad() : nullptr);
        sw->backwardDone_ = true;
        inputGradDone_ = true;
        return;
      }
    }
    static const bool enabled = std::getenv("MRN_NO_GROUPED_DX") == nullptr;
    if(!enabled)
      return;
    std::vector<AffineNodeOp*> group{this};
    for(auto& u : upcoming) {
      auto* a = dynamic_cast<AffineNodeOp*>(u.get());
      if(!a || !a->adj_ || a->child(0) != child(0) || !a->trainable() || a->inputGradDone_ || a->shape() != shape() || a->child(1)->shape() != child(1)->shape())
        break;
      group.push_back(a);
    }
    if(group.size() < 2)
      return;
    std::vector<Tensor> adjs, weights, biasGrads;
    bool allBias = true;
    for(auto* a : group) {
      adjs.push_back(a->adj_);
      weights.push_back(a->child(1)->val());
      allBias = allBias && a->fuseBias();
    }
    if(allBias)
      for(auto* a : group)
        biasGrads.push_back(a->child(2)->grad());
    static const bool trace = std::getenv("MRN_PEEPHOLE_TRACE") != nullptr;
    if(trace)
      fprintf(stderr, "[peephole] grouped input gradient of %d projections of node %zu\n", (int)group.size(), (size_t)child(0)->getId());
    ProdGroupedNT(getBackend()->getGemmHandle(), child(0)->grad(), adjs, weights, 1.0, biasGrads);
    for(auto* a : group) {
      a->inputGradDone_ = true;
      a->biasGradDone_ = allBias;
    }
    // ... and their weight gradients dW_g += X^T adj_g as one launch as well (off the critical path)
    bool allWeights = true;
    std::vector<Tensor> weightGrads;
    for(auto* a : group) {
      allWeights = allWeights && a->child(1)->trainable() && a->child(1)->type() == "param" && !a->weightGradDone_;
      if(allWeights)
        weightGrads.push_back(a->child(1)->grad());
    }
    if(allWeights) {
      bool done = false;
      offCriticalPath(child(1), [&] { done = ProdSharedA(getBackend()->getGemmHandle(), weightGrads, child(0)->val(), adjs, {}, true, 1.f); });
      if(done)
        for(auto* a : group)
          a->weightGradDone_ = true;
    }
  }
  NodeOps backwardOps() {
    using namespace functional;
    // dW and db hang off the backward chain: side stream when W / b are parameters
    return {NodeOp(offCriticalPath(child(0), [&] {
              if(inputGradDone_)
                return;
              if(fuseBias()) {  // single projection: same product, bias gradient from its A tiles
                ProdGroupedNT(getBackend()->getGemmHandle(), child(0)->grad(), {adj_}, {child(1)->val()}, 1.0, {child(2)->grad()});
                biasGradDone_ = true;
              } else {
                Prod(getBackend()->getGemmHandle(), child(0)->grad(), adj_, child(1)->val(), false, true, 1.0);
              }
            })),
            NodeOp(offCriticalPath(child(1), [&] {
              if(!weightGradDone_)
                Prod(getBackend()->getGemmHandle(), child(1)->grad(), child(0)->val(), adj_, true, false, 1.0);
            })),
            NodeOp(offCriticalPath(child(2), [&] {
              if(!biasGradDone_) {
                ABORT_IF(adj_->memory()->fp32Skipped, "affine: bias gradient needs the fp32 adjoint, but only its bf16 copy was written");
                Add(_1, child(2)->grad(), adj_);
              }
              // column sums the input-gradient product queued instead of taking them from its own tiles: issued here,
              // behind this node's weight-gradient product on the side stream
              ProdFlushColumnSums(getBackend()->getGemmHandle());
            }))};
  }
  const std::string type() { return "affine"; }
};

// sum_axis(a * b).  reference: :371-404
struct ScalarProductNodeOp : public NaryNodeOp {
  ScalarProductNodeOp(Expr a, Expr b, keywords::axis_k ax) : NaryNodeOp({a, b}, newShape(a, b, ax.value)) {}
  static Shape newShape(Expr a, Expr b, int ax) {
    Shape full = Shape::broadcast({a, b});
    full.set(full.axis(ax), 1);
    return full;
  }
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Reduce(_1 * _2, val_, child(0)->val(), child(1)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1 * _2, child(0)->grad(), child(1)->val(), adj_)),
            NodeOp(Add(_1 * _2, child(1)->grad(), child(0)->val(), adj_))};
  }
  const std::string type() { return "scalar-product"; }
};

struct ElementBinaryNodeOp : public NaryNodeOp {
  ElementBinaryNodeOp(Expr a, Expr b) : NaryNodeOp({a, b}, Shape::broadcast({a, b})) {}
};

// reference: :418-436
struct PlusNodeOp : public ElementBinaryNodeOp {
  PlusNodeOp(Expr a, Expr b) : ElementBinaryNodeOp(a, b) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = _2 + _3, val_, child(0)->val(), child(1)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1, child(0)->grad(), adj_)), NodeOp(Add(_1, child(1)->grad(), adj_))};
  }
  const std::string type() { return "+"; }
};

// reference: :438-456
struct MinusNodeOp : public ElementBinaryNodeOp {
  MinusNodeOp(Expr a, Expr b) : ElementBinaryNodeOp(a, b) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = _2 - _3, val_, child(0)->val(), child(1)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1, child(0)->grad(), adj_)), NodeOp(Add(-_1, child(1)->grad(), adj_))};
  }
  const std::string type() { return "-"; }
};

// reference: :458-476
struct MultNodeOp : public ElementBinaryNodeOp {
  MultNodeOp(Expr a, Expr b) : ElementBinaryNodeOp(a, b) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = _2 * _3, val_, child(0)->val(), child(1)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1 * _2, child(0)->grad(), adj_, child(1)->val())),
            NodeOp(Add(_1 * _2, child(1)->grad(), adj_, child(0)->val()))};
  }
  const std::string type() { return "x"; }
};

// reference: :478-505
struct DivNodeOp : public ElementBinaryNodeOp {
  DivNodeOp(Expr a, Expr b) : ElementBinaryNodeOp(a, b) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = _2 / _3, val_, child(0)->val(), child(1)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1 * 1.0f / _2, child(0)->grad(), adj_, child(1)->val())),
            NodeOp(Add(-_1 * _2 / (_3 * _3), child(1)->grad(), adj_, child(0)->val(), child(1)->val()))};
  }
  const std::string type() { return "/"; }
};

// -log softmax(a)[b] per row; labels are a float tensor.  reference: :531-553
struct CrossEntropyNodeOp : public NaryNodeOp {
  CrossEntropyNodeOp(Expr a, Expr b) : NaryNodeOp({a, b}, newShape(a)) {}
  ~CrossEntropyNodeOp() {
    auto g = graph();
    if(stats_ && g)
      g->free(stats_);
  }
  static Shape newShape(Expr a) {
    Shape shape1 = a->shape();
    shape1.set(-1, 1);
    return shape1;
  }
  // the forward kernel leaves (max, sum exp) per row for the backward kernel
  void forward() {
    if(!stats_)
      graph()->tensor(stats_, Shape{(int)val_->size(), 2});
    CrossEntropyPick(val_, child(0)->val(), child(1)->val(), stats_);
  }
  NodeOps backwardOps() {
    return {NodeOp(CrossEntropyPickBackward(child(0)->grad(), adj_, child(0)->val(), child(1)->val(), stats_))};
  }
  const std::string type() { return "x-ent"; }

private:
  Tensor stats_;
};

// reference: :555-613.  backward() zeroes the child adjoints itself and
// Deconcatenate ASSIGNS (kept as is).
struct ConcatenateNodeOp : public NaryNodeOp {
  ConcatenateNodeOp(const std::vector<Expr>& nodes, keywords::axis_k ax)
      : NaryNodeOp(nodes, newShape(nodes, ax.value)), ax_(nodes.back()->shape().axis(ax.value)) {}

  static Shape newShape(const std::vector<Expr>& nodes, int ax) {
    Shape shape = nodes.back()->shape();
    int a = shape.axis(ax);
    int sum = 0;
    for(auto child : nodes)
      sum += child->shape()[a];
    shape.set(a, sum);
    return shape;
  }

  void forward() {
    std::vector<Tensor> concatenees;
    for(size_t i = 0; i < children_.size(); ++i)
      concatenees.push_back(child(i)->val());
    Concatenate(val_, concatenees, ax_);
  }
  void backward() {
    // [U | Ux]-style concatenations of parameters: the weight-gradient products that fill adj_ run on the side
    // stream (Node::offCriticalPath), so this node's backward follows them there
    offCriticalPath(shared_from_this(), [&] {
      std::vector<Tensor> deconcatenees;
      for(size_t i = 0; i < children_.size(); ++i) {
        auto childPtr = child(i);
        childPtr->set_zero_adjoint();
        deconcatenees.push_back(childPtr->grad());
      }
      Deconcatenate(deconcatenees, adj_, ax_);
    });
  }
  virtual bool paramOnly() {
    if(paramOnly_ < 0) {
      paramOnly_ = 1;
      for(auto& c : children_)
        if(!c->paramOnly())
          paramOnly_ = 0;
    }
    return paramOnly_ == 1;
  }
  int paramOnly_{-1};

  virtual size_t hash() {
    if(!hash_) {
      hash_ = NaryNodeOp::hash();
      hash_combine(hash_, ax_);
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<ConcatenateNodeOp>(node);
    return cnode && ax_ == cnode->ax_;
  }
  const std::string type() { return "concat"; }
  int ax_;
};

// reference: :657-688.  ONE backward closure guarded by child(0) (the input);
// gamma/beta gradients are produced by the same kernel.
struct LayerNormalizationOp : public NaryNodeOp {
  LayerNormalizationOp(const std::vector<Expr>& nodes, float eps = 1e-9) : NaryNodeOp(nodes), eps_(eps) {}

  NodeOps forwardOps() {
    return {NodeOp(LayerNormalization(
        val_, child(0)->val(), child(1)->val(), (children_.size() == 3) ? child(2)->val() : nullptr, eps_))};
  }
  NodeOps backwardOps() {
    return {NodeOp(LayerNormalizationGrad(child(0)->grad(),
                                          child(1)->grad(),
                                          (children_.size() == 3) ? child(2)->grad() : nullptr,
                                          adj_,
                                          val_,
                                          child(0)->val(),
                                          child(1)->val(),
                                          (children_.size() == 3) ? child(2)->val() : nullptr,
                                          eps_))};
  }
  virtual size_t hash() {
    if(!hash_) {
      hash_ = NaryNodeOp::hash();
      hash_combine(hash_, eps_);
    }
    return hash_;
  }
  const std::string type() { return "layer_normalization"; }

private:
  float eps_;
};

// layer_norm(x + residual): children {x, residual, gamma, beta}.  The reference builds a
// PlusNodeOp and a LayerNormalizationOp (Transformer::PostProcess "a" then "n",
// src/models/transformer.h:111-123); fused here the sum is never written to memory and the
// backward pass delivers d(x + r) to both inputs from one kernel.
struct ResidualLayerNormOp : public NaryNodeOp {
  ResidualLayerNormOp(const std::vector<Expr>& nodes, float eps) : NaryNodeOp(nodes), eps_(eps) {
    ABORT_IF(nodes.size() != 4, "residual layer-norm expects {x, residual, gamma, beta}");
    ABORT_IF(nodes[0]->shape() != nodes[1]->shape(), "residual layer-norm: x and residual must have the same shape");
  }
  NodeOps forwardOps() {
    return {NodeOp(LayerNormalization(val_, child(0)->val(), child(2)->val(), child(3)->val(), eps_, child(1)->val()))};
  }
  void backward() {
    bool wantX = child(0)->trainable(), wantR = child(1)->trainable();
    ABORT_IF(!wantX, "residual layer-norm expects a trainable main input");
    LayerNormalizationGrad(child(0)->grad(), child(2)->grad(), child(3)->grad(), adj_, val_, child(0)->val(), child(2)->val(), child(3)->val(), eps_,
                           child(1)->val(), wantR ? child(1)->grad() : Tensor());
  }
  virtual size_t hash() {
    if(!hash_) {
      hash_ = NaryNodeOp::hash();
      hash_combine(hash_, eps_);
    }
    return hash_;
  }
  const std::string type() { return "residual_layer_normalization"; }

private:
  float eps_;
};

// Fused multi-head attention core: children {q, k, v[, additive mask]} in [beam, B, T, d]
// layout.  Replaces the node sequence SplitHeads x3 / bdot / + / softmax / bdot / JoinHeads of
// the reference's Transformer::MultiHead (src/models/transformer.h:153-261) by one forward and
// one backward kernel; the softmax output is kept in a workspace tensor for the backward pass.
struct MultiHeadAttentionNodeOp : public NaryNodeOp {
  MultiHeadAttentionNodeOp(const std::vector<Expr>& nodes, int heads, float scale)
      : NaryNodeOp(nodes, nodes[0]->shape()), heads_(heads), scale_(scale) {
    if(nodes.size() > 3)
      ABORT_IF(nodes[3]->trainable(), "attention mask must not require a gradient");
  }
  ~MultiHeadAttentionNodeOp() {
    auto g = graph();
    if(probs_ && g)
      g->free(probs_);
  }

  void forward() {
    int Tq = child(0)->shape()[-2], Tk = child(1)->shape()[-2];
    int batch = child(0)->shape().elements() / (Tq * child(0)->shape()[-1]);
    if(!probs_)
      graph()->tensor(probs_, Shape{batch, heads_, Tq, Tk});
    MultiHeadAttention(val_, probs_, child(0)->val(), child(1)->val(), child(2)->val(), children_.size() > 3 ? child(3)->val() : nullptr, heads_, scale_, exact());
  }
  void backward() {
    // q, k, v come out of trainable projections in every model; a frozen input would still
    // get a (discarded) adjoint from set_zero_adjoint only if trainable, so require it
    ABORT_IF(!child(0)->trainable() || !child(1)->trainable() || !child(2)->trainable(), "fused attention expects trainable q, k, v");
    MultiHeadAttentionGrad(child(0)->grad(), child(1)->grad(), child(2)->grad(), adj_, val_, probs_, child(0)->val(), child(1)->val(), child(2)->val(), heads_, scale_, exact());
  }
  // fp32-grade products in the exact GEMM modes, plain tf32 in the tf32 / bf16 modes
  bool exact() {
    auto mode = getGemmMode(getBackend()->getGemmHandle());
    return mode == GemmMode::FP32 || mode == GemmMode::BF16X3;
  }

  virtual size_t hash() {
    if(!hash_) {
      hash_ = NaryNodeOp::hash();
      hash_combine(hash_, heads_);
      hash_combine(hash_, scale_);
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<MultiHeadAttentionNodeOp>(node);
    return cnode && heads_ == cnode->heads_ && scale_ == cnode->scale_;
  }
  const std::string type() { return "multi-head-attention"; }

private:
  int heads_;
  float scale_;
  Tensor probs_;
};

// sigma(t)*y + (1-sigma(t))*x.  reference: :690-709 (backward assigns)
struct HighwayNodeOp : public NaryNodeOp {
  HighwayNodeOp(const std::vector<Expr>& nodes) : NaryNodeOp(nodes) {}
  NodeOps forwardOps() {
    return {NodeOp(HighwayForward(val_, child(0)->val(), child(1)->val(), child(2)->val()))};
  }
  NodeOps backwardOps() {
    return {NodeOp(HighwayBackward(
        child(0)->grad(), child(1)->grad(), child(2)->grad(), child(0)->val(), child(1)->val(), child(2)->val(), adj_))};
  }
  const std::string type() { return "highway"; }
};

}  // namespace marian
