// Single-input graph nodes.  Forward/backward formulas follow the reference's
// src/graph/node_operators_unary.h (cited per class); each closure is one or
// two tensor-operator calls (kernels/tensor_operators.h).
#pragma once

#include "graph/backend.h"
#include "graph/expression_graph.h"
#include "graph/node.h"
#include "kernels/tensor_operators.h"

namespace marian {

struct UnaryNodeOp : public NaryNodeOp {
  UnaryNodeOp(Expr a, const Shape& shape) : NaryNodeOp({a}, shape) {}
  explicit UnaryNodeOp(Expr a) : NaryNodeOp({a}, a->shape()) {}
};

// Helper: a unary node with one float attribute that takes part in CSE.
template <class Derived>
struct ScalarAttrNodeOp : public UnaryNodeOp {
  float scalar_{0};
  ScalarAttrNodeOp(Expr a, float scalar) : UnaryNodeOp(a), scalar_(scalar) {}
  virtual size_t hash() {
    if(!hash_) {
      hash_ = NaryNodeOp::hash();
      hash_combine(hash_, scalar_);
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<Derived>(node);
    return cnode && scalar_ == cnode->scalar_;
  }
};

// reference: node_operators_unary.h:22-60
struct ScalarAddNodeOp : public ScalarAttrNodeOp<ScalarAddNodeOp> {
  ScalarAddNodeOp(Expr a, float scalar) : ScalarAttrNodeOp(a, scalar) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = _2 + scalar_, val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1, child(0)->grad(), adj_))};
  }
  const std::string type() { return "scalar_add"; }
};

// reference: :62-102 (its type() string is also "scalar_add"; CSE still tells the
// two apart there through dynamic_cast in equal(); we give it its own name)
struct ScalarMultNodeOp : public ScalarAttrNodeOp<ScalarMultNodeOp> {
  ScalarMultNodeOp(Expr a, float scalar) : ScalarAttrNodeOp(a, scalar) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = scalar_ * _2, val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(scalar_ * _1, child(0)->grad(), adj_))};
  }
  const std::string type() { return "scalar_mult"; }
};

// reference: :104-118
struct LogitNodeOp : public UnaryNodeOp {
  LogitNodeOp(Expr a) : UnaryNodeOp(a) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = logit(_2), val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1 * _2 * (1.0f - _2), child(0)->grad(), adj_, val_))};
  }
  const std::string type() { return "logit"; }
};

// tanh(a [+ b [+ c ...]])  reference: :163-211
struct TanhNodeOp : public NaryNodeOp {
  TanhNodeOp(const std::vector<Expr>& nodes) : NaryNodeOp(nodes, Shape::broadcast(nodes)) {}

  NodeOps forwardOps() {
    using namespace functional;
    switch(children_.size()) {
      case 1: return {NodeOp(Element(_1 = tanh(_2), val_, child(0)->val()))};
      case 2: return {NodeOp(Element(_1 = tanh(_2 + _3), val_, child(0)->val(), child(1)->val()))};
      case 3:
        return {NodeOp(Element(_1 = tanh(_2 + _3 + _4), val_, child(0)->val(), child(1)->val(), child(2)->val()))};
      default:
        return {NodeOp(Element(_1 = _2 + _3 + _4, val_, child(0)->val(), child(1)->val(), child(2)->val());
                       for(size_t i = 3; i < children_.size(); ++i) Element(_1 = _1 + _2, val_, child(i)->val());
                       Element(_1 = tanh(_1), val_);)};
    }
  }
  NodeOps backwardOps() {
    using namespace functional;
    NodeOps ops;
    for(size_t i = 0; i < children_.size(); i++)
      ops.push_back(NodeOp(Add(_1 * (1.0f - (_2 * _2)), child(i)->grad(), adj_, val_)));
    return ops;
  }
  const std::string type() { return "tanh"; }
};

// reference: :228-248
struct ReLUNodeOp : public UnaryNodeOp {
  ReLUNodeOp(Expr a) : UnaryNodeOp(a) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = ReLU(_2), val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1 * ReLUback(_2), child(0)->grad(), adj_, child(0)->val()))};
  }
  const std::string type() { return "ReLU"; }
};

// reference: :279-314  (note its constructor order: alpha first)
struct PReLUNodeOp : public ScalarAttrNodeOp<PReLUNodeOp> {
  PReLUNodeOp(float alpha, Expr a) : ScalarAttrNodeOp(a, alpha) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = PReLU(_2, scalar_), val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1 * PReLUback(_2, scalar_), child(0)->grad(), adj_, child(0)->val()))};
  }
  const std::string type() { return "PReLU"; }
};

// x * sigmoid(x)   reference: :330-351
struct SwishNodeOp : public UnaryNodeOp {
  SwishNodeOp(Expr a) : UnaryNodeOp(a) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = _2 * logit(_2), val_, child(0)->val()))};
  }
  // set by the consumer that delivered the gradient of the pre-activation itself
  // (AffineNodeOp::fuseBackward: swish' applied in the epilogue of its input-gradient product)
  bool backwardDone_{false};
  // BF16S GEMM mode: when every consumer of swish(x) is a product, only its bf16 copy is written (the fp32 tensor
  // is half of the feed-forward block's forward write traffic and nobody reads it); the unfused backward pass then
  // recomputes f(x) = x sigma(x) from x instead of reading it.
  size_t allocate() {
    size_t n = Node::allocate();
    if(val_ && wantValShadow_ && consumers_ > 0 && productConsumers_ == consumers_ && (shape_[-1] % 8) == 0)
      val_->memory()->shadowOnly = true;
    return n;
  }
  NodeOps backwardOps() {
    using namespace functional;
    // dJ/dx += dJ/df * (f(x) + sigma(x) * (1 - f(x)))
    return {NodeOp(if(!backwardDone_) {
      if(val_->memory()->fp32Skipped)
        Add(_1 * (_2 * logit(_2) + logit(_2) * (1.f - _2 * logit(_2))), child(0)->grad(), adj_, child(0)->val());
      else
        Add(_1 * (_3 + logit(_2) * (1.f - _3)), child(0)->grad(), adj_, child(0)->val(), val_);
    })};
  }
  const std::string type() { return "swish"; }
};

// reference: :353-408.  The mask is NOT a child (no gradient flows to it).
struct SoftmaxNodeOp : public NaryNodeOp {
  Expr mask_;
  SoftmaxNodeOp(Expr a, Expr mask = nullptr) : NaryNodeOp({a}, a->shape()), mask_(mask) {}

  NodeOps forwardOps() { return {NodeOp(Softmax(val_, child(0)->val(), mask_ ? mask_->val() : nullptr))}; }
  NodeOps backwardOps() { return {NodeOp(SoftmaxGrad(child(0)->grad(), adj_, val_))}; }

  virtual size_t hash() {
    if(!hash_) {
      hash_ = NaryNodeOp::hash();
      if(mask_)
        hash_combine(hash_, mask_->hash());
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<SoftmaxNodeOp>(node);
    if(!cnode)
      return false;
    if((bool)mask_ != (bool)cnode->mask_)
      return false;
    if(mask_ && !mask_->equal(cnode->mask_))
      return false;
    return true;
  }
  const std::string type() { return "softmax"; }
};

// reference: :410-424
struct LogSoftmaxNodeOp : public UnaryNodeOp {
  LogSoftmaxNodeOp(Expr a) : UnaryNodeOp(a) {}
  NodeOps forwardOps() { return {NodeOp(LogSoftmax(val_, child(0)->val()))}; }
  NodeOps backwardOps() { return {NodeOp(LogSoftmaxGrad(child(0)->grad(), adj_, val_))}; }
  const std::string type() { return "logsoftmax"; }
};

// Helper for nodes with one int attribute (axis).
template <class Derived>
struct AxisNodeOp : public UnaryNodeOp {
  int ax_;
  AxisNodeOp(Expr a, int ax) : UnaryNodeOp(a, reducedShape(a, ax)), ax_(a->shape().axis(ax)) {}
  static Shape reducedShape(Expr a, int ax) {
    Shape shape = a->shape();
    shape.set(shape.axis(ax), 1);
    return shape;
  }
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
    auto cnode = std::dynamic_pointer_cast<Derived>(node);
    return cnode && ax_ == cnode->ax_;
  }
};

// reference: :426-472
struct SumNodeOp : public AxisNodeOp<SumNodeOp> {
  SumNodeOp(Expr a, keywords::axis_k ax) : AxisNodeOp(a, ax.value) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Reduce(_1, val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1, child(0)->grad(), adj_))};
  }
  const std::string type() { return "sum"; }
};

// reference: :474-528
struct MeanNodeOp : public AxisNodeOp<MeanNodeOp> {
  MeanNodeOp(Expr a, keywords::axis_k ax) : AxisNodeOp(a, ax.value) {}
  NodeOps forwardOps() {
    using namespace functional;
    int left = child(0)->shape().elements() / val_->shape().elements();
    float scale = 1.f / left;
    return {NodeOp(Reduce(_1, scale, val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    int left = child(0)->shape().elements() / val_->shape().elements();
    float scale = 1.f / left;
    return {NodeOp(Add(_1, scale, child(0)->grad(), adj_))};
  }
  const std::string type() { return "mean"; }
};

// reference: :530-546
struct LogNodeOp : public UnaryNodeOp {
  LogNodeOp(Expr a) : UnaryNodeOp(a) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = log(_2), val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1 * (1.f / _2), child(0)->grad(), adj_, child(0)->val()))};
  }
  const std::string type() { return "log"; }
};

// reference: :548-563
struct ExpNodeOp : public UnaryNodeOp {
  ExpNodeOp(Expr a) : UnaryNodeOp(a) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = exp(_2), val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(_1 * exp(_2), child(0)->grad(), adj_, child(0)->val()))};
  }
  const std::string type() { return "exp"; }
};

// sqrt(x + eps)   reference: :565-603
struct SqrtNodeOp : public ScalarAttrNodeOp<SqrtNodeOp> {
  SqrtNodeOp(Expr a, float epsilon) : ScalarAttrNodeOp(a, epsilon) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = sqrt(_2 + scalar_), val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(0.5f * (1.f / _1) * _2, child(0)->grad(), val_, adj_))};
  }
  const std::string type() { return "sqrt"; }
};

// reference: :605-621
struct SquareNodeOp : public UnaryNodeOp {
  SquareNodeOp(Expr a) : UnaryNodeOp(a) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = _2 * _2, val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(2.f * _1 * _2, child(0)->grad(), child(0)->val(), adj_))};
  }
  const std::string type() { return "square"; }
};

// reference: :623-638
struct NegNodeOp : public UnaryNodeOp {
  NegNodeOp(Expr a) : UnaryNodeOp(a) {}
  NodeOps forwardOps() {
    using namespace functional;
    return {NodeOp(Element(_1 = -_2, val_, child(0)->val()))};
  }
  NodeOps backwardOps() {
    using namespace functional;
    return {NodeOp(Add(-_1, child(0)->grad(), adj_))};
  }
  const std::string type() { return "-"; }
};

// Embedding lookup: out[i,:] = a[idx[i],:]; backward scatter-adds.
// reference: :640-691 (RowsNodeOp) -> CopyRows / PasteRows.  The index vector is
// uploaded once to workspace memory as int32 (see ExpressionGraph::uploadIndices);
// when it comes from the batch, a refill closure makes the upload replayable.
struct RowsNodeOp : public UnaryNodeOp {
  RowsNodeOp(Expr a,
             const std::vector<size_t>& indices,
             ExpressionGraph::BatchFillI fill = nullptr,
             Ptr<data::CorpusBatch> batch = nullptr)
      : UnaryNodeOp(a, newShape(a, indices)), indices_(indices), fill_(fill), batch_(batch) {}

  ~RowsNodeOp() {
    auto g = graph();
    if(g && devIdx_)
      g->allocator()->free(devIdx_);
  }

  const int* deviceIndices() {
    if(!devIdx_) {
      if(fill_)
        devIdx_ = graph()->uploadIndices(indices_.size(), fill_, batch_);
      else
        devIdx_ = graph()->uploadIndices(indices_);
    }
    return (const int*)devIdx_->data();
  }

  NodeOps forwardOps() { return {NodeOp(CopyRows(val_, child(0)->val(), deviceIndices(), indices_.size()))}; }
  NodeOps backwardOps() { return {NodeOp(PasteRows(child(0)->grad(), adj_, deviceIndices(), indices_.size()))}; }

  static Shape newShape(Expr a, const std::vector<size_t>& indices) {
    Shape shape = a->shape();
    ABORT_IF(shape.size() != 2, "rows operator can only be used with 2-dimensional tensors");
    shape.set(0, (int)indices.size());
    return shape;
  }

  // A lookup whose indices are refilled from the batch on every replay of a captured step (fill_) is its own node:
  // two such lookups with equal capture-time indices (copy task, tied source / target embeddings) read DIFFERENT
  // batch streams afterwards, merging them would feed one stream to both.
  virtual size_t hash() {
    if(!hash_) {
      size_t seed = NaryNodeOp::hash();
      for(auto i : indices_)
        hash_combine(seed, i);
      if(fill_)
        hash_combine(seed, (size_t)this);
      hash_ = seed;
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(fill_)
      return this == node.get();
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<RowsNodeOp>(node);
    return cnode && !cnode->fill_ && indices_ == cnode->indices_;
  }
  const std::string type() { return "rows"; }

  std::vector<size_t> indices_;
  ExpressionGraph::BatchFillI fill_;
  Ptr<data::CorpusBatch> batch_;
  Ptr<MemoryPiece> devIdx_;
};

// reference: :693-742
struct ColsNodeOp : public UnaryNodeOp {
  ColsNodeOp(Expr a, const std::vector<size_t>& indices) : UnaryNodeOp(a, newShape(a, indices)), indices_(indices) {}
  NodeOps forwardOps() { return {NodeOp(CopyCols(val_, child(0)->val(), indices_))}; }
  NodeOps backwardOps() { return {NodeOp(PasteCols(child(0)->grad(), adj_, indices_))}; }
  static Shape newShape(Expr a, const std::vector<size_t>& indices) {
    Shape shape = a->shape();
    shape.set(1, (int)indices.size());
    return shape;
  }
  virtual size_t hash() {
    if(!hash_) {
      size_t seed = NaryNodeOp::hash();
      for(auto i : indices_)
        hash_combine(seed, i);
      hash_ = seed;
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<ColsNodeOp>(node);
    return cnode && indices_ == cnode->indices_;
  }
  const std::string type() { return "cols"; }
  std::vector<size_t> indices_;
};

// reference: :744-796
struct SelectNodeOp : public UnaryNodeOp {
  SelectNodeOp(Expr a, int axis, const std::vector<size_t>& indices)
      : UnaryNodeOp(a, newShape(a, axis, indices)), indices_(indices), axis_(a->shape().axis(axis)) {}
  NodeOps forwardOps() { return {NodeOp(Select(graph()->allocator(), val_, child(0)->val(), axis_, indices_))}; }
  NodeOps backwardOps() { return {NodeOp(Insert(graph()->allocator(), child(0)->grad(), adj_, axis_, indices_))}; }
  static Shape newShape(Expr a, int axis, const std::vector<size_t>& indices) {
    Shape shape = a->shape();
    shape.set(shape.axis(axis), (int)indices.size());
    return shape;
  }
  virtual size_t hash() {
    if(!hash_) {
      size_t seed = NaryNodeOp::hash();
      hash_combine(seed, axis_);
      for(auto i : indices_)
        hash_combine(seed, i);
      hash_ = seed;
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<SelectNodeOp>(node);
    return cnode && axis_ == cnode->axis_ && indices_ == cnode->indices_;
  }
  const std::string type() { return "select"; }
  std::vector<size_t> indices_;
  int axis_{0};
};

// reference: :798-851.  NB: backward ASSIGNS into the child gradient (TransposeND
// overwrites), it does not accumulate - kept for results parity.
struct TransposeNodeOp : public UnaryNodeOp {
  std::vector<int> axes_;
  TransposeNodeOp(Expr a, const std::vector<int>& axes) : UnaryNodeOp(a, newShape(a, axes)), axes_{axes} {}
  NodeOps forwardOps() { return {NodeOp(TransposeND(val_, child(0)->val(), axes_))}; }
  NodeOps backwardOps() { return {NodeOp(TransposeND(child(0)->grad(), adj_, axes_))}; }
  static Shape newShape(Expr a, const std::vector<int>& axes) {
    Shape shape = a->shape();
    ABORT_IF(shape.size() != axes.size(), "Shape and transpose axes have different number of dimensions");
    for(size_t i = 0; i < shape.size(); ++i)
      shape.set((int)i, a->shape()[axes[i]]);
    return shape;
  }
  virtual size_t hash() {
    if(!hash_) {
      size_t seed = NaryNodeOp::hash();
      for(auto ax : axes_)
        hash_combine(seed, ax);
      hash_ = seed;
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<TransposeNodeOp>(node);
    return cnode && axes_ == cnode->axes_;
  }
  const std::string type() { return "transpose"; }
};

// Zero-copy view with another shape.  reference: :853-912
class ReshapeNodeOp : public UnaryNodeOp {
private:
  Expr reshapee_;

public:
  ReshapeNodeOp(Expr a, Shape shape) : UnaryNodeOp(a, shape), reshapee_(a) {
    ABORT_IF(shape.elements() != a->shape().elements(), "reshape changes the number of elements");
    Node::destroy_ = false;
  }
  ~ReshapeNodeOp() {}

  size_t allocate() { return 0; }
  void free() {}
  void forward() {}
  void backward() {}
  void init_dependent() { reshapee_->init_dependent(); }
  void set_zero_adjoint() { reshapee_->set_zero_adjoint(); }
  // value and adjoint alias the reshapee's: shadow requests and consumer counts belong there
  void requestValShadow() { reshapee_->requestValShadow(); }
  void addConsumer(bool viaProduct = false) { reshapee_->addConsumer(viaProduct); }
  bool isView() const { return true; }
  // lanes: the adjoint IS the reshapee's - writers and waiters see through the view
  void noteConsumerLane(int l) {
    Node::noteConsumerLane(l);
    reshapee_->noteConsumerLane(l);
  }
  void setAdjMark(int l, void* m) {
    Node::setAdjMark(l, m);
    reshapee_->setAdjMark(l, m);
  }
  void waitAdjMarks() {
    Node::waitAdjMarks();
    reshapee_->waitAdjMarks();
  }

  Tensor& val() {
    auto childVal = reshapee_->val();
    val_.reset(new TensorBase(childVal->memory(), shape(), childVal->getDevice()));
    return val_;
  }
  Tensor& grad() {
    auto childGrad = reshapee_->grad();
    adj_.reset(new TensorBase(childGrad->memory(), shape(), childGrad->getDevice()));
    return adj_;
  }

  virtual size_t hash() {
    if(!hash_) {
      size_t seed = NaryNodeOp::hash();
      for(auto s : shape())
        hash_combine(seed, s);
      hash_ = seed;
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<ReshapeNodeOp>(node);
    return cnode && shape() == cnode->shape();
  }
  const std::string type() { return "reshape"; }
};

// Zero-copy view of time step `step` along `axis`.  reference: :914-993
class StepNodeOp : public UnaryNodeOp {
private:
  Expr stepNode_;
  int step_;
  int axis_;

public:
  StepNodeOp(Expr a, int step, int axis)
      : UnaryNodeOp(a, newShape(a, axis)), stepNode_(a), step_(step), axis_(a->shape().axis(axis)) {
    Node::destroy_ = false;
  }

  static Shape newShape(Expr a, int axis) {
    Shape outShape = a->shape();
    int ax = outShape.axis(axis);
    for(int i = 0; i <= ax; ++i)
      outShape.set(i, 1);
    return outShape;
  }

  size_t allocate() { return 0; }
  void free() {}
  void forward() {}
  void backward() {}
  void init_dependent() { stepNode_->init_dependent(); }
  void set_zero_adjoint() { stepNode_->set_zero_adjoint(); }
  // lanes: the adjoint is a slice of the stepped node's - writers and waiters see through the view
  void noteConsumerLane(int l) {
    Node::noteConsumerLane(l);
    stepNode_->noteConsumerLane(l);
  }
  void setAdjMark(int l, void* m) {
    Node::setAdjMark(l, m);
    stepNode_->setAdjMark(l, m);
  }
  void waitAdjMarks() {
    Node::waitAdjMarks();
    stepNode_->waitAdjMarks();
  }

  Tensor& val() {
    auto childVal = stepNode_->val();
    size_t offset = (size_t)step_ * shape().elements() * sizeof(float);
    auto mem = New<MemoryPiece>(childVal->memory()->data() + offset, childVal->memory()->size());
    val_.reset(new TensorBase(mem, shape(), childVal->getDevice()));
    return val_;
  }
  Tensor& grad() {
    auto childGrad = stepNode_->grad();
    size_t offset = (size_t)step_ * shape().elements() * sizeof(float);
    // data(): a partial view cannot carry the lazy-zero state of the whole adjoint
    auto mem = New<MemoryPiece>((uint8_t*)childGrad->data() + offset, childGrad->memory()->size());
    adj_.reset(new TensorBase(mem, shape(), childGrad->getDevice()));
    return adj_;
  }

  virtual size_t hash() {
    if(!hash_) {
      hash_ = NaryNodeOp::hash();
      hash_combine(hash_, step_);
      hash_combine(hash_, axis_);
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<StepNodeOp>(node);
    return cnode && step_ == cnode->step_ && axis_ == cnode->axis_;
  }
  const std::string type() { return "step"; }
};

// reference: :995-1032; backward assigns (Shift with invert=true).
struct ShiftNodeOp : public UnaryNodeOp {
  ShiftNodeOp(Expr a, Shape shift) : UnaryNodeOp(a), shift_(shift) {}
  NodeOps forwardOps() { return {NodeOp(Shift(val_, child(0)->val(), shift_))}; }
  NodeOps backwardOps() { return {NodeOp(Shift(child(0)->grad(), adj_, shift_, true))}; }
  virtual size_t hash() {
    if(!hash_) {
      size_t seed = NaryNodeOp::hash();
      for(auto i : shift_)
        hash_combine(seed, i);
      hash_ = seed;
    }
    return hash_;
  }
  virtual bool equal(Expr node) {
    if(!NaryNodeOp::equal(node))
      return false;
    auto cnode = std::dynamic_pointer_cast<ShiftNodeOp>(node);
    return cnode && shift_ == cnode->shift_;
  }
  const std::string type() { return "shift"; }
  Shape shift_;
};

}  // namespace marian
