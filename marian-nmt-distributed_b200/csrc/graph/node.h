// Chainable / Node / NaryNodeOp: one element of the define-by-run tape.
//
// API-compatible with the reference (src/graph/chainable.h:59-99,
// src/graph/node.h:14-190): forwardOps()/backwardOps() return closures calling
// the tensor operators; runBackward() skips non-trainable children; nodes are
// hashed for common-subexpression elimination by (name, type, children ids
// [+ op attributes]).  The boost::hash machinery is replaced by a local
// hash_combine; graphviz/debug plumbing is reduced to what tests use.
#pragma once

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

#include "common/definitions.h"
#include "common/keywords.h"
#include "common/shape.h"
#include "tensors/device.h"
#include "tensors/tensor.h"

namespace marian {

#define NodeOp(op) [=]() { op; }
typedef std::vector<std::function<void()>> NodeOps;

class ExpressionGraph;
class Backend;

template <class T>
inline void hash_combine(size_t& seed, const T& v) {
  seed ^= std::hash<T>()(v) + 0x9e3779b97f4a7c15ULL + (seed << 6) + (seed >> 2);
}

template <class DataType>
struct Chainable {
  Chainable() {}
  virtual ~Chainable() {}

  virtual void forward() = 0;
  virtual void backward() = 0;
  virtual void fuseBackward(const std::vector<std::shared_ptr<Chainable<DataType>>>& /*upcoming*/) {}
  virtual void fuseForward(const std::vector<std::shared_ptr<Chainable<DataType>>>& /*upcoming*/) {}
  virtual NodeOps forwardOps() = 0;
  virtual NodeOps backwardOps() = 0;

  virtual size_t allocate() = 0;
  virtual void free() = 0;
  virtual void init() = 0;
  virtual void init_dependent() {}
  virtual void set_zero_adjoint() {}
  virtual bool trainable() = 0;
  virtual void setTrainable(bool) = 0;

  virtual void setId(size_t) = 0;
  virtual size_t getId() = 0;

  virtual Ptr<ExpressionGraph> graph() = 0;
  virtual const Shape& shape() = 0;
  virtual std::vector<Expr>& children() = 0;
  virtual Expr child(size_t) = 0;
  virtual DataType& val() = 0;
  virtual DataType& grad() = 0;
  virtual float scalar() = 0;

  virtual const std::string type() = 0;
  virtual void set_name(const std::string&) = 0;
  virtual const std::string& name() const = 0;

  virtual void debug(const std::string& message) = 0;
  virtual bool marked_for_debug() = 0;
  virtual const std::string& debug_message() = 0;

  virtual size_t hash() = 0;
  virtual bool equal(Expr) = 0;

  // scheduling hints (see Node): forward pass may overlap with the main stream
  virtual void setConcurrent(bool = true) {}
  virtual bool concurrent() const { return false; }
  virtual bool sideProduced() const { return false; }
  virtual void setSideProduced(bool) {}

  // lanes (tensors/device.h, ExpressionGraph::setLane): which independent chain of the pass this node belongs to, and
  // the bookkeeping that orders its value / adjoint against nodes of other lanes
  virtual int lane() const { return 0; }
  virtual void setLane(int) {}
  virtual void noteConsumerLane(int) {}
  virtual bool crossLane() const { return false; }
  virtual bool adjSharedAcrossLanes(int /*writerLane*/) const { return false; }
  virtual void* valMark() const { return nullptr; }
  virtual void setValMark(void*) {}
  virtual void setAdjMark(int /*lane*/, void*) {}
  virtual void waitAdjMarks() {}

  // bf16 shadows of GEMM operands (kernels/shadow.h, GemmMode::BF16S).  A product node asks its
  // operand nodes for a bf16 copy of their VALUE and is itself marked as wanting one of its ADJOINT
  // (the A / B operand of its two backward products).  The graph counts the consumers of every node
  // when they are added to the tape: an adjoint shadow is only requested for nodes with exactly one
  // consumer, i.e. one writer of the adjoint.  Views forward all three to the node they alias.
  virtual void requestValShadow() {}
  // `viaProduct`: the consumer reads this node's value only as a tensor-core product operand (through the bf16 copy)
  virtual void addConsumer(bool /*viaProduct*/ = false) {}
  virtual bool isView() const { return false; }
  // true for parameters and for nodes built from parameters only (the [U | Ux] concatenations of the recurrent
  // cells): nothing in the backward sweep reads their adjoints before the optimizer, so whatever writes them may
  // leave the critical path (Node::offCriticalPath)
  virtual bool paramOnly() { return false; }
  // child i of this node is read only as a product operand (Affine / Dot nodes: their two matrix arguments)
  virtual bool readsChildViaProduct(size_t /*i*/) const { return false; }
};

class Node : public Chainable<Tensor>, public std::enable_shared_from_this<Node> {
protected:
  size_t id_{0};
  bool trainable_{true};
  bool destroy_{true};
  std::vector<Expr> children_;
  Weak<ExpressionGraph> graph_;
  Shape shape_{1, 1, 1, 1};
  std::string name_{"none"};
  Tensor val_{nullptr};
  Tensor adj_{nullptr};
  bool markedForDebug_{false};
  std::string debugMessage_;
  bool concurrent_{false};      // forward pass may run on the side stream (see setConcurrent)
  bool sideProduced_{false};    // val_ was written on the side stream and not yet joined
  bool wantValShadow_{false};   // a product reads val_: producers leave a bf16 copy (BF16S GEMM mode)
  bool wantAdjShadow_{false};   // this node is a product: its adjoint is an operand of the backward products
  int consumers_{0};            // nodes on the tape that have this node as a child
  int productConsumers_{0};     // ... of which read the value only as a product operand
  int lane_{0};                 // see Chainable::lane
  unsigned consumerLanes_{0};   // bit k: a consumer lives on lane k
  void* valMark_{nullptr};      // device::laneMark behind this node's forward (only taken for cross-lane consumers)
  void* adjMarks_[device::kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};  // per lane: mark behind its latest write to adj_

public:
  Node(Ptr<ExpressionGraph> graph, const Shape& shape) : graph_(graph), shape_(shape) {}

  virtual ~Node() {
    if(destroy_)
      free();
  }

  virtual float scalar();

  virtual NodeOps forwardOps() { return {}; }
  virtual NodeOps backwardOps() { return {}; }

  virtual void runForward(const NodeOps& ops) {
    for(auto&& op : ops)
      op();
  }
  // Backward closure i belongs to child i and only runs if that child needs a
  // gradient (reference: node.h:57-62).
  virtual void runBackward(const NodeOps& ops) {
    size_t i = 0;
    for(auto&& op : ops)
      if(child(i++)->trainable())
        op();
  }

  virtual void forward() { runForward(forwardOps()); }
  virtual void backward() { runBackward(backwardOps()); }
  // Peephole of the backward sweep: `upcoming` holds the nodes that run right after this one (next
  // first).  Their adjoints are complete (all their consumers precede them in the sweep and this
  // node is none of them), so a node may take over part of their backward work and mark them.
  virtual void fuseBackward(const std::vector<Expr>& /*upcoming*/) {}
  // Peephole of the forward pass: `upcoming` holds the nodes that run right after this one (next first); a node may
  // compute their values together with its own (after allocating them) and mark them done.
  virtual void fuseForward(const std::vector<Expr>& /*upcoming*/) {}

  virtual bool trainable() { return trainable_; }
  virtual void setTrainable(bool trainable) { trainable_ = trainable; }

  virtual void setId(size_t id) { id_ = id; }
  virtual size_t getId() { return id_; }

  virtual Ptr<ExpressionGraph> graph() { return graph_.lock(); }

  virtual void debug(const std::string& message) {
    debugMessage_ = message;
    markedForDebug_ = true;
  }
  virtual bool marked_for_debug() { return markedForDebug_; }
  virtual const std::string& debug_message() { return debugMessage_; }

  virtual size_t allocate();
  virtual void free();
  virtual void init() {}
  virtual void init_dependent();
  virtual void set_zero_adjoint();

  virtual Tensor& val() { return val_; }
  virtual Tensor& grad() { return adj_; }
  virtual const Shape& shape() { return shape_; }

  void set_name(const std::string& name) { name_ = name; }
  const std::string& name() const { return name_; }

  virtual std::vector<Expr>& children() { return children_; }
  virtual Expr child(size_t i) { return children_[i]; }

  // Scheduling hint from model code: this node's forward only depends on values that exist when
  // the previously created node has run, and its result is not needed immediately - e.g. the key
  // and value projections of an attention block while the query projection runs.  The graph runs
  // it on the side stream and joins before the first consumer (ExpressionGraph::forwardNext).
  virtual void setConcurrent(bool c = true) { concurrent_ = c; }
  virtual bool concurrent() const { return concurrent_; }
  virtual bool sideProduced() const { return sideProduced_; }
  virtual void setSideProduced(bool s) { sideProduced_ = s; }

  virtual void requestValShadow() {
    wantValShadow_ = true;
    if(val_)
      val_->memory()->shadowWanted = true;
  }
  virtual void addConsumer(bool viaProduct = false) {
    ++consumers_;
    if(viaProduct)
      ++productConsumers_;
  }

  virtual int lane() const { return lane_; }
  virtual void setLane(int l) { lane_ = l; }
  virtual void noteConsumerLane(int l) { consumerLanes_ |= 1u << l; }
  // some consumer reads this node's value from another lane
  virtual bool crossLane() const { return (consumerLanes_ & ~(1u << lane_)) != 0; }
  // the adjoint is written by consumers (on their lanes) and read by this node's backward (on its own): does any of
  // those parties live on a lane other than `writerLane`?
  virtual bool adjSharedAcrossLanes(int writerLane) const { return ((consumerLanes_ | (1u << lane_)) & ~(1u << writerLane)) != 0; }
  virtual void* valMark() const { return valMark_; }
  virtual void setValMark(void* m) { valMark_ = m; }
  virtual void setAdjMark(int l, void* m) { adjMarks_[l] = m; }
  // the current lane waits for every other lane's latest write to this node's adjoint
  virtual void waitAdjMarks() {
    for(int l = 0; l < device::kMaxLanes; ++l)
      if(adjMarks_[l])
        device::laneWait(adjMarks_[l]);
  }

  Ptr<Backend> getBackend();

  // Runs f on the side stream when it only produces the gradient of PARAMETER `target` (or of a node made of
  // parameters only, whose own backward then runs on the side stream as well): nothing downstream in the
  // backward sweep reads it (device.h: forkSide/joinSide).
  template <class F>
  void offCriticalPath(Expr target, F f) {
    static const bool enabled = std::getenv("MRN_NO_SIDE_STREAM") == nullptr;
    bool side = enabled && target->paramOnly();
    if(side)
      device::forkSide();
    f();
    if(side)
      device::returnFromSide();
  }
};

struct NaryNodeOp : public Node {
  size_t hash_{0};

  NaryNodeOp(const std::vector<Expr>& nodes, const Shape& shape) : Node(nodes.front()->graph(), shape) {
    setup(nodes);
  }
  // default shape: the first child's
  explicit NaryNodeOp(const std::vector<Expr>& nodes) : Node(nodes.front()->graph(), nodes.front()->shape()) {
    setup(nodes);
  }

  virtual ~NaryNodeOp() {}

  virtual size_t hash() {
    if(!hash_) {
      size_t seed = std::hash<std::string>()(name());
      hash_combine(seed, type());
      for(size_t i = 0; i < children_.size(); ++i)
        hash_combine(seed, child(i)->hash());
      hash_ = seed;
    }
    return hash_;
  }

  virtual bool equal(Expr node) {
    if(type() != node->type())
      return false;
    if(name() != node->name())
      return false;
    if(children().size() != node->children().size())
      return false;
    for(size_t i = 0; i < children().size(); ++i)
      if(children()[i]->getId() != node->children()[i]->getId())
        return false;
    return true;
  }

private:
  void setup(const std::vector<Expr>& nodes);
};

}  // namespace marian
