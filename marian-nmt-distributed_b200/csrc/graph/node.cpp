// Node memory management + graph-side upload helpers.
// Reference behaviour: src/graph/node.cu:7-51 (allocate / free / init_dependent
// / set_zero_adjoint), src/graph/expression_graph.cu:24-35 (dropout mask node).
#include <cstdlib>

#include "graph/node.h"
#include "graph/expression_graph.h"
#include "data/batch.h"

namespace marian {

size_t Config::seed = 1234;

Staging*& currentStagingSlot() {
  static thread_local Staging* slot = nullptr;
  return slot;
}

size_t Node::allocate() {
  size_t elements = 0;
  if(!val_) {
    graph()->tensor(val_, shape_);
    elements = val_->shape().elements();
    if(wantValShadow_)
      val_->memory()->shadowWanted = true;
  }
  return elements;
}

void Node::free() {
  auto g = graph();
  if(g) {
    if(val_)
      g->free(val_);
    if(adj_)
      g->free(adj_);
  }
}

void Node::init_dependent() {
  if(!adj_) {
    graph()->tensor(adj_, shape_);
    adj_->set(1);
  }
}

void Node::set_zero_adjoint() {
  if(!adj_) {
    graph()->tensor(adj_, shape_);
    // zeroed on first touch, or assigned by the first accumulating writer
    // (MRN_EAGER_ZERO=1 restores the reference's memset-then-accumulate for debugging)
    static const bool eager = std::getenv("MRN_EAGER_ZERO") != nullptr;
    static const char* eagerTypes = std::getenv("MRN_EAGER_TYPES");
    if(eager || (eagerTypes && std::string(eagerTypes).find("," + type() + ",") != std::string::npos))
      adj_->set(0);
    else
      adj_->setLazyZero();
    // one consumer = one writer: that writer may leave the bf16 copy the backward products read
    if(wantAdjShadow_ && consumers_ == 1)
      adj_->memory()->shadowWanted = true;
  }
}

float Node::scalar() {
  return val_->scalar();
}

Ptr<Backend> Node::getBackend() {
  return graph()->getBackend();
}

void NaryNodeOp::setup(const std::vector<Expr>& nodes) {
  children_.resize(nodes.size());
  for(size_t i = 0; i < nodes.size(); ++i)
    children_[i] = nodes[i];
  setTrainable(std::any_of(nodes.begin(), nodes.end(), [](Expr a) { return a->trainable(); }));
  for(auto child : children_)
    graph()->remove_top_node(child);
}

Expr ExpressionGraph::batchConstant(Shape shape, BatchFillF fill, Ptr<data::CorpusBatch> batch) {
  auto self = shared_from_this();
  size_t n = shape.elements();
  auto init = [self, fill, batch, n](Tensor t) {
    float* pinned = (float*)self->staging().take(n * sizeof(float));
    fill(*batch, pinned);
    device::copyH2D(t->data(), pinned, n * sizeof(float));
    self->batchUploads().push_back(BatchUpload{
        pinned, n * sizeof(float), [fill](void* p, const data::CorpusBatch& b) { fill(b, (float*)p); }});
  };
  return constant(shape, keywords::init = std::function<void(Tensor)>(init));
}

Ptr<MemoryPiece> ExpressionGraph::uploadIndices(const std::vector<size_t>& indices) {
  size_t n = indices.size();
  auto mem = allocator()->alloc<int>(n);
  int* pinned = (int*)staging_->take(n * sizeof(int));
  for(size_t i = 0; i < n; ++i)
    pinned[i] = (int)indices[i];
  device::copyH2D(mem->data(), pinned, n * sizeof(int));
  return mem;
}

Ptr<MemoryPiece> ExpressionGraph::uploadIndices(size_t n, BatchFillI fill, Ptr<data::CorpusBatch> batch) {
  auto mem = allocator()->alloc<int>(n);
  int* pinned = (int*)staging_->take(n * sizeof(int));
  fill(*batch, pinned);
  device::copyH2D(mem->data(), pinned, n * sizeof(int));
  batchUploads_.push_back(
      BatchUpload{pinned, n * sizeof(int), [fill](void* p, const data::CorpusBatch& b) { fill(b, (int*)p); }});
  return mem;
}

Expr ExpressionGraph::dropout(float prob, Shape shape) {
  auto backend = backend_;
  auto init = [prob, backend](Tensor t) { Dropout(t, prob, backend->nextDropoutSeed(), backend->dropoutEpoch()); };
  return constant(shape, keywords::init = std::function<void(Tensor)>(init));
}

}  // namespace marian
