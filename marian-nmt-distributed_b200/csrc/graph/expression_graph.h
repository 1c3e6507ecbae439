// ExpressionGraph: the define-by-run tape + parameter store of one device.
//
// API and execution order follow the reference (src/graph/expression_graph.h:28-512,
// src/graph/parameters.h:10-89, src/graph/node_operators.h:8-80):
//   add()      CSE by hash, tape push, top-node bookkeeping          (:385-409)
//   forward()  params arena, then allocate/init/forward per node      (:134-162)
//   backward() zero param grads, seed top node with 1, reverse sweep,
//              zero child adjoints lazily, drop children as we go     (:180-215)
// Parameters live in two exact-size arenas (values / gradients) that expose the
// whole model as ONE flat tensor - the layout the optimizer and the multi-GPU
// shard exchange rely on.
//
// B200-first differences: nothing here synchronises the stream; constants and
// index vectors are uploaded through pinned staging owned by the graph; batch
// dependent uploads are recorded (batchUploads()) so a captured CUDA graph of
// the whole step can be replayed on the next batch of the same shape
// (training/graph_replay.h).
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <map>
#include <unordered_map>
#include <unordered_set>

#include "common/definitions.h"
#include "common/keywords.h"
#include "graph/backend.h"
#include "graph/node.h"
#include "kernels/tensor_operators.h"
#include "layers/param_initializers.h"
#include "tensors/allocator.h"
#include "tensors/staging.h"

namespace marian {

namespace data {
class CorpusBatch;
}

struct ConstantNode : public Node {
  ConstantNode(Ptr<ExpressionGraph> graph, const Shape& shape, std::function<void(Tensor)> init)
      : Node(graph, shape), init_(init) {
    setTrainable(false);
  }
  virtual void init() {
    if(!initialized_) {
      init_(val_);
      initialized_ = true;
    }
  }
  const std::string type() { return "const"; }
  virtual size_t hash() {
    size_t seed = std::hash<std::string>()(name());
    hash_combine(seed, type());
    hash_combine(seed, (size_t)this);
    return seed;
  }
  virtual bool equal(Expr node) { return this == node.get(); }

protected:
  std::function<void(Tensor)> init_;
  bool initialized_{false};
};

struct ParamNode : public Node {
  ParamNode(Ptr<ExpressionGraph> graph, const Shape& shape, std::function<void(Tensor)> init, bool fixed)
      : Node(graph, shape), init_(init) {
    setTrainable(!fixed);
  }
  virtual void init() {
    if(!initialized_) {
      init_(val_);
      initialized_ = true;
    }
  }
  const std::string type() { return "param"; }
  virtual bool paramOnly() { return true; }
  virtual size_t hash() {
    size_t seed = std::hash<std::string>()(name());
    hash_combine(seed, type());
    hash_combine(seed, (size_t)this);
    return seed;
  }
  virtual bool equal(Expr node) { return name() == node->name(); }

private:
  std::function<void(Tensor)> init_;
  bool initialized_{false};
};

class Parameters {
public:
  void init(int device) {
    vals_ = New<TensorAllocator>(device);
    grads_ = New<TensorAllocator>(device);
  }

  std::vector<Expr>::iterator begin() { return params_.begin(); }
  std::vector<Expr>::iterator end() { return params_.end(); }
  std::map<std::string, Expr>& getMap() { return named_; }

  Expr get(const std::string& name) {
    auto it = named_.find(name);
    return it != named_.end() ? it->second : Expr();
  }
  size_t size() { return params_.size(); }

  // Bytes of one flat arena.  With setShardCount(N) the arena is padded so it
  // splits into N equal, 256-byte aligned shards (what reduce-scatter /
  // all-gather need); the padding holds zeros and never receives gradients.
  size_t totalCapacity(Ptr<TensorAllocator> alloc) {
    size_t sum = 0;
    for(auto p : params_)
      sum += alloc->capacity(p->shape());
    size_t quantum = 256 * (size_t)shards_;
    return (sum + quantum - 1) / quantum * quantum;
  }
  void setShardCount(int n) { shards_ = n > 0 ? n : 1; }
  int shardCount() const { return shards_; }

  void add(Expr p, const std::string& name) {
    ABORT_IF(named_.count(name), "Parameter already exists:", name);
    params_.push_back(p);
    named_[name] = p;
  }

  void allocateForward() {
    if(vals_->size() == 0 && !params_.empty()) {
      vals_->reserveExact(totalCapacity(vals_));
      vals_->asTensor()->set(0);  // padding must be defined (flat tensor is exchanged as a whole)
      for(auto p : params_)
        if(!p->val())
          vals_->allocate(p->val(), p->shape());
    }
  }
  void allocateBackward() {
    if(grads_->size() == 0 && !params_.empty()) {
      grads_->reserveExact(totalCapacity(grads_));
      for(auto p : params_)
        if(!p->grad())
          grads_->allocate(p->grad(), p->shape());
    }
  }
  void set_zero_adjoint() { grads()->set(0); }

  Tensor vals() { return vals_->asTensor(); }
  Tensor grads() { return grads_->asTensor(); }

  void clear() {
    params_.clear();
    named_.clear();
    vals_->clear();
    grads_->clear();
  }

private:
  std::vector<Expr> params_;
  std::map<std::string, Expr> named_;
  Ptr<TensorAllocator> vals_;
  Ptr<TensorAllocator> grads_;
  int shards_{1};
};

template <class T, typename... Args>
Expr Expression(Args&&... args);

class ExpressionGraph : public std::enable_shared_from_this<ExpressionGraph> {
public:
  explicit ExpressionGraph(bool inference = false) : inferenceOnly_(inference) {}
  ExpressionGraph(const ExpressionGraph&) = delete;

  ~ExpressionGraph() {
    clear();
    if(params_)
      params_->clear();
  }

  void setInference(bool inference) { inferenceOnly_ = inference; }
  bool inference() const { return inferenceOnly_; }

  void setDevice(size_t device = 0) {
    device_ = (int)device;
    device::setDevice(device_);
    params_ = New<Parameters>();
    params_->init(device_);
    tensors_ = New<TensorAllocator>(device_);
    backend_ = New<Backend>(device_, Config::seed);
  }
  size_t getDevice() { return device_; }
  Ptr<Backend> getBackend() { return backend_; }

  void switchParams(const std::string& newNamespace) { namespace_ = newNamespace; }

  void reserveWorkspaceMB(size_t num) { tensors_->reserve(num * 1024 * 1024 - 1); }

  void copyParams(Ptr<ExpressionGraph> graph) {
    for(auto p : *graph->params())
      param(p->name(), p->shape());
    params()->allocateForward();
    params()->vals()->copyFrom(graph->params()->vals());
  }

  void backprop() {
    forward();
    backward();
  }

  bool fits() {
    try {
      tensors_->throwAtReallocation(true);
      backprop();
      tensors_->throwAtReallocation(false);
    } catch(AllocationException&) {
      tensors_->throwAtReallocation(false);
      return false;
    }
    return true;
  }

  void forward() {
    device::setDevice(device_);
    params_->allocateForward();
    backend_->newForwardPass();
    gemmInvalidateCache(backend_->getGemmHandle());
    if(params_->size() > 0) {
      auto vals = params_->vals();
      gemmSetStableRange(backend_->getGemmHandle(), vals->data(), vals->size() * sizeof(float));
    }
    forwardNext();
  }

  void forwardNext() {
    StagingScope scope(staging_.get());
    hashMap_.clear();
    static const bool sideEnabled = std::getenv("MRN_NO_SIDE_STREAM") == nullptr;
    bool pending = false;  // side-stream results not yet joined into the main stream
    // (an inference pass frees values as it goes - memory would be reused across lanes without any ordering)
    const bool lanes = lanesInUse_ && !inferenceOnly_ && !nodeTiming();
    if(lanes)
      device::openLanes();
    while(!nodesForward_.empty()) {
      auto v = nodesForward_.front();
      if(lanes) {
        // this node's chain; values produced by other chains are waited for one by one
        device::selectLane(v->lane());
        for(auto& c : v->children())
          if(c->lane() != v->lane()) {
            if(!c->valMark()) {  // (no mark taken behind the producer: everything its lane has issued so far)
              device::selectLane(c->lane());
              c->setValMark(device::laneMark());
              device::selectLane(v->lane());
            }
            device::laneWait(c->valMark());
          }
      }
      v->allocate();
      v->init();
      // (not in an inference pass: values are freed as soon as their last consumer has been ISSUED - a value read only
      // by side-stream nodes, e.g. the per-beam view of the encoder context under the key / value projections of a
      // decoding step, would be recycled by the main stream while the side stream still reads it)
      if(sideEnabled && !lanes && !inferenceOnly_ && v->concurrent() && !v->children().empty()) {
        // inputs were produced on the main stream before this point (or on the side stream
        // itself, which is in order): fork, run, hand back
        device::forkSide();
        v->forward();
        device::returnFromSide();
        v->setSideProduced(true);
        pending = true;
      } else {
        if(pending) {
          bool fromSide = false;
          for(auto& c : v->children())
            fromSide = fromSide || c->sideProduced();
          if(fromSide) {
            const std::string t = v->type();
            if(t == "reshape" || t == "step") {
              v->setSideProduced(true);  // zero-copy view: no device work, the dependency moves on
            } else {
              device::joinSide();  // joins ALL side work
              pending = false;
              for(auto& c : v->children())
                c->setSideProduced(false);
            }
          }
        }
        if(!pending && !lanes) {  // (a node computed together with its successors must not race with side-stream producers)
          std::vector<Expr> upcoming;
          auto it = nodesForward_.begin();
          for(++it; it != nodesForward_.end() && upcoming.size() < 2; ++it)
            upcoming.push_back(*it);
          v->fuseForward(upcoming);
        }
        if(nodeTiming()) {
          auto t0 = std::chrono::steady_clock::now();
          v->forward();
          timing_["fwd " + v->type()] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        } else {
          v->forward();
        }
      }
      if(lanes && v->crossLane())
        v->setValMark(device::laneMark());
      if(inferenceOnly_)
        v->children().clear();
      nodesForward_.pop_front();
    }
    if(pending)
      device::joinSide();
    if(lanes)
      device::closeLanes();
  }

  void backward() {
    ABORT_IF(topNodes_.size() > 1, "There are more than one top most node for backward step");
    device::setDevice(device_);
    StagingScope scope(staging_.get());

    params_->allocateBackward();
    params_->set_zero_adjoint();

    // Weight/bias gradients are issued on the side stream (Node::offCriticalPath) and may still
    // be reading a node's adjoint after the node is gone: frees are parked until the join.
    tensors_->allocator()->deferFrees(true);

    for(auto&& v : topNodes_)
      v->init_dependent();

    topNodes_.clear();
    hashMap_.clear();

    // optional split of the sweep (training/graph_group.h: gradient exchange overlapped with the rest of the sweep)
    size_t splitAfter = (size_t)-1, swept = 0;
    if(backwardSplitChooser_)
      splitAfter = backwardSplitChooser_(nodesBackward_);

    const bool lanes = lanesInUse_ && !nodeTiming();
    if(lanes)
      device::openLanes();
    while(!nodesBackward_.empty()) {
      if(swept++ == splitAfter && backwardSplitHook_) {
        ProdFlushColumnSums(backend_->getGemmHandle());
        if(lanes)
          device::closeLanes();
        device::joinSide();  // weight gradients issued so far are part of "before the split"
        backwardSplitHook_();
        if(lanes)
          device::openLanes();
      }
      auto v = nodesBackward_.back();
      nodesBackward_.pop_back();
      if(lanes)
        device::selectLane(v->lane());

      {
        auto tz = std::chrono::steady_clock::now();
        for(auto&& child : v->children())
          if(child->trainable())
            child->set_zero_adjoint();
        if(nodeTiming())
          timing_["(zero adjoints)"] += std::chrono::duration<double>(std::chrono::steady_clock::now() - tz).count();
      }

      if(v->trainable()) {
        if(lanes) {
          // this node's chain waits for the other chains' writes to the adjoints it reads (its own) and updates
          // (its children's: accumulation is read-modify-write)
          v->waitAdjMarks();
          for(auto&& child : v->children())
            if(child->trainable())
              child->waitAdjMarks();
        } else {
          std::vector<Expr> upcoming;
          for(auto it = nodesBackward_.rbegin(); it != nodesBackward_.rend() && upcoming.size() < 2; ++it)
            upcoming.push_back(*it);
          v->fuseBackward(upcoming);
        }
        if(nodeTiming()) {
          auto t0 = std::chrono::steady_clock::now();
          v->backward();
          timing_["bwd " + v->type()] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        } else {
          v->backward();
        }
        if(lanes) {
          void* mark = nullptr;
          for(auto&& child : v->children())
            if(child->trainable() && child->adjSharedAcrossLanes(v->lane())) {
              if(!mark)
                mark = device::laneMark();
              child->setAdjMark(v->lane(), mark);
            }
        }
      }

      {
        auto tz = std::chrono::steady_clock::now();
        v->children().clear();
        if(nodeTiming())
          timing_["(release children)"] += std::chrono::duration<double>(std::chrono::steady_clock::now() - tz).count();
      }
    }
    ProdFlushColumnSums(backend_->getGemmHandle());  // (a bias without gradient closure: nothing may stay queued)
    if(lanes)
      device::closeLanes();
    if(nodeTiming()) {
      // host wall time per node type: meaningful on the synchronous CPU oracle (MRN_NODE_TIMING=1)
      std::vector<std::pair<double, std::string>> rows;
      for(auto& kv : timing_)
        rows.push_back({kv.second, kv.first});
      std::sort(rows.rbegin(), rows.rend());
      for(size_t i = 0; i < rows.size() && i < 14; ++i)
        fprintf(stderr, "[node-timing] %-40s %8.3f s\n", rows[i].second.c_str(), rows[i].first);
      timing_.clear();
    }
    device::joinSide();
    tensors_->allocator()->deferFrees(false);
    tensors_->allocator()->flushDeferred();
  }

  template <typename... Args>
  Expr param(std::string name, Shape shape, Args... args) {
    if(!namespace_.empty())
      name = namespace_ + "::" + name;

    bool fixed = keywords::Get(keywords::fixed, false, args...);
    auto p = params_->get(name);
    if(p) {
      ABORT_IF(shape != p->shape(), "Requested shape for existing parameter does not match original shape:", name);
      p->setTrainable(!fixed);
      add(p);
      return p;
    }
    ABORT_IF(reloaded_, "Graph was reloaded and parameter is newly created:", name);

    std::function<void(Tensor)> init = keywords::Get(keywords::init, std::function<void(Tensor)>([](Tensor) {}), args...);
    p = Expression<ParamNode>(shared_from_this(), shape, init, fixed);
    p->set_name(name);
    params_->add(p, name);
    return p;
  }

  template <typename... Args>
  Expr constant(Shape shape, Args... args) {
    std::function<void(Tensor)> init = keywords::Get(keywords::init, std::function<void(Tensor)>([](Tensor) {}), args...);
    return Expression<ConstantNode>(shared_from_this(), shape, init);
  }
  Expr ones(Shape shape) { return constant(shape, keywords::init = inits::ones); }
  Expr zeros(Shape shape) { return constant(shape, keywords::init = inits::zeros); }

  // Constant whose content is a function of the current batch (indices, masks).
  // `fill` writes shape.elements() floats into pinned staging; it is called now
  // and again on every replay of a captured step with the next batch.
  typedef std::function<void(const data::CorpusBatch&, float*)> BatchFillF;
  typedef std::function<void(const data::CorpusBatch&, int*)> BatchFillI;
  Expr batchConstant(Shape shape, BatchFillF fill, Ptr<data::CorpusBatch> batch);

  // Device int32 index vector (embedding rows) uploaded through staging.  The
  // returned piece is workspace memory; the caller frees it via allocator().
  Ptr<MemoryPiece> uploadIndices(const std::vector<size_t>& indices);
  Ptr<MemoryPiece> uploadIndices(size_t n, BatchFillI fill, Ptr<data::CorpusBatch> batch);

  Expr dropout(float prob, Shape shape);

  Expr get(std::string name) {
    if(!namespace_.empty())
      name = namespace_ + "::" + name;
    return params_->get(name);
  }

  Ptr<Parameters>& params() { return params_; }

  // Lanes: nodes created while lane k is selected form a chain of their own; forward and backward run chain k > 0 on
  // its own stream, ordered against the other chains only where values / adjoints actually cross (tensors/device.h).
  // Model code brackets an independent sub-network: graph->setLane(1); ... ; graph->setLane(0);
  void setLane(int lane) { currentLane_ = (lanesAllowed() && !inferenceOnly_ && lane > 0 && lane < device::kMaxLanes) ? lane : 0; }
  int lane() const { return currentLane_; }
  static bool lanesAllowed() {
    static const bool on = std::getenv("MRN_NO_LANES") == nullptr;
    return on;
  }

  // Backward split: `chooser` sees the tape (forward order; the sweep runs it back to front) and returns after how
  // many swept nodes `hook` is to be called (or size_t(-1): never).  The side stream is joined before the hook.
  typedef std::function<size_t(const std::list<Expr>&)> SplitChooser;
  void setBackwardSplit(SplitChooser chooser, std::function<void()> hook) {
    backwardSplitChooser_ = chooser;
    backwardSplitHook_ = hook;
  }

  Expr add(Expr node) {
    size_t hash = node->hash();
    auto it = hashMap_.find(hash);
    if(it != hashMap_.end()) {
      for(auto foundWeak : it->second) {
        auto found = foundWeak.lock();
        if(found && node->equal(found))
          return found;
      }
    }
    hashMap_[hash].push_back(node);
    node->setId(count_++);
    if(!node->isView()) {  // a view hands its consumers through to the node it aliases
      size_t i = 0;
      for(auto& child : node->children())
        child->addConsumer(node->readsChildViaProduct(i++));
    }
    node->setLane(currentLane_);
    if(currentLane_ != 0)
      lanesInUse_ = true;
    for(auto& child : node->children())
      child->noteConsumerLane(currentLane_);

    nodesForward_.push_back(node);
    if(!inferenceOnly_ && node->trainable()) {
      nodesBackward_.push_back(node);
      topNodes_.insert(node);
    }
    return node;
  }

  void remove_top_node(Expr node) { topNodes_.erase(node); }

  void tensor(Tensor& t, const Shape& shape) { tensors_->allocate(t, shape); }
  void free(Tensor& t) {
    if(tensors_)
      tensors_->free(t);
  }
  Ptr<Allocator> allocator() { return tensors_->allocator(); }

  void clear() {
    count_ = 0;
    currentLane_ = 0;
    lanesInUse_ = false;
    nodesForward_.clear();
    nodesBackward_.clear();
    topNodes_.clear();
    hashMap_.clear();
    if(tensors_)
      tensors_->clear();
    batchUploads_.clear();
    staging_->reset();
  }

  void clearParameters() { params_->clear(); }
  void setReloaded(bool reloaded) { reloaded_ = reloaded; }

  // --- replay support --------------------------------------------------
  std::vector<BatchUpload>& batchUploads() { return batchUploads_; }
  Staging& staging() { return *staging_; }
  // Hands the pinned staging (and the recorded uploads) of the current tape to
  // a captured step; the graph continues with a fresh staging.
  Ptr<Staging> detachStaging() {
    auto s = staging_;
    staging_ = New<Staging>();
    return s;
  }
  size_t numNodes() const { return count_; }

  struct StagingScope {
    Staging* prev;
    explicit StagingScope(Staging* s) : prev(currentStagingSlot()) { currentStagingSlot() = s; }
    ~StagingScope() { currentStagingSlot() = prev; }
  };

private:
  static bool nodeTiming() {
    static const bool on = std::getenv("MRN_NODE_TIMING") != nullptr;
    return on;
  }
  std::map<std::string, double> timing_;
  size_t count_{0};
  int currentLane_{0};
  bool lanesInUse_{false};  // some node of the current tape lives on a lane > 0
  std::list<Expr> nodesForward_;
  std::list<Expr> nodesBackward_;
  SplitChooser backwardSplitChooser_;
  std::function<void()> backwardSplitHook_;
  std::unordered_set<Expr> topNodes_;
  Ptr<Parameters> params_;
  Ptr<TensorAllocator> tensors_;
  int device_{0};
  Ptr<Backend> backend_;
  std::unordered_map<size_t, std::vector<WExpr>> hashMap_;
  bool inferenceOnly_{false};
  bool reloaded_{false};
  std::string namespace_;
  Ptr<Staging> staging_{New<Staging>()};
  std::vector<BatchUpload> batchUploads_;
};

template <class T, typename... Args>
Expr Expression(Args&&... args) {
  auto e = Expr(new T(std::forward<Args>(args)...));
  return e->graph()->add(e);
}

}  // namespace marian
