// Whole-step CUDA-graph capture and replay.
//
// The reference rebuilds the tape and launches ~1.5k kernels (plus hundreds of
// stream syncs) for every batch (SURVEY.md section 7 "hard parts").  The tape
// topology and every buffer address only depend on the batch SHAPE: the model
// code is deterministic, the workspace arena never moves and is reset before
// each build.  So the first time a shape is seen after parameters exist, the
// forward+backward sweep is recorded into a CUDA graph; afterwards a step is
//   refill pinned staging with the new batch  ->  one cudaGraphLaunch.
// The define-by-run API is untouched: a new shape simply builds a new tape.
//
// What is recorded: constant/index uploads (as memcpy nodes reading the
// plan-owned pinned staging), all kernels of forward() and backward(), and the
// D2H copy of the cost scalar.  The optimizer step stays outside (its bias
// correction terms change every step).
#pragma once

#include <map>
#include <vector>

#include "data/batch.h"
#include "graph/expression_graph.h"
#include "tensors/device.h"

namespace marian {

class StepReplay {
public:
  struct Plan {
    void* exec{nullptr};
    Ptr<Staging> staging;
    std::vector<BatchUpload> uploads;
    size_t launches{0};
    size_t kernels{0};  // kernel nodes recorded in the graph
    void* done{nullptr};  // marker behind the plan's latest launch: its upload nodes have read the pinned staging
    // split step (gradient exchange overlapped with the backward sweep): `exec` ends at the split of the sweep,
    // `exec2` is the rest; `midTag` is what the group decided at the split (restored before its hook runs)
    void* exec2{nullptr};
    int midTag{-1};
  };

  // what runs between the two halves of a split step
  struct MidStep {
    virtual ~MidStep() {}
    virtual size_t choose(const std::list<Expr>& tape) = 0;  // after how many swept nodes (size_t(-1): no split); remembers the decision
    virtual int tag() = 0;
    virtual void restore(int tag) = 0;
    virtual void run() = 0;
  };

  ~StepReplay() { clear(); }

  void clear() {
    for(auto& it : plans_) {
      if(it.second.exec)
        device::destroyGraph(it.second.exec);
      if(it.second.exec2)
        device::destroyGraph(it.second.exec2);
      device::freeMarker(it.second.done);
    }
    plans_.clear();
    seen_.clear();
  }

  bool enabled() const { return enabled_ && device::captureSupported(); }
  void setEnabled(bool e) { enabled_ = e; }

  Plan* find(const std::vector<int>& key) {
    auto it = plans_.find(key);
    return it == plans_.end() ? nullptr : &it->second;
  }

  // A shape is captured the SECOND time it shows up: the first pass runs
  // eagerly so that parameter initialisation, arena growth and packed-weight
  // scratch growth all happen outside a recording.
  bool shouldCapture(const std::vector<int>& key) {
    if(!enabled())
      return false;
    return seen_[key]++ >= 1;
  }

  // The graph's memcpy nodes read the plan's pinned staging when they EXECUTE: the host must not refill it for the
  // next batch while an earlier launch of the same plan is still queued (the host runs several steps ahead of the
  // device when nobody reads the cost back).  The optimizer step that follows each launch keeps the device busy
  // while the host waits here, so this costs no device time.
  void replay(Plan& plan, const data::CorpusBatch& batch, MidStep* mid = nullptr) {
    device::waitMarker(plan.done);
    for(auto& u : plan.uploads)
      u.refill(u.pinned, batch);
    launch(plan, mid);
  }
  void launch(Plan& plan, MidStep* mid = nullptr) {
    device::launchGraph(plan.exec);
    if(plan.exec2) {
      ABORT_IF(!mid, "a split step needs its mid-step hook");
      mid->restore(plan.midTag);
      mid->run();
      device::launchGraph(plan.exec2);
    }
    plan.done = device::recordMarker(plan.done);
    plan.launches++;
  }

  // Takes ownership of the recorded graph and of the tape's staging/uploads.
  Plan& store(const std::vector<int>& key, void* exec, Ptr<ExpressionGraph> graph, void* exec2 = nullptr, int midTag = -1, size_t kernelsFirst = 0) {
    Plan& p = plans_[key];
    p.exec = exec;
    p.exec2 = exec2;
    p.midTag = midTag;
    p.kernels = device::lastCaptureKernelCount() + kernelsFirst;
    lastKernels_ = p.kernels;
    p.uploads = graph->batchUploads();
    p.staging = graph->detachStaging();
    return p;
  }

  size_t size() const { return plans_.size(); }
  size_t lastPlanKernels() const { return lastKernels_; }

private:
  bool enabled_{true};
  size_t lastKernels_{0};
  std::map<std::vector<int>, Plan> plans_;
  std::map<std::vector<int>, int> seen_;
};

}  // namespace marian
