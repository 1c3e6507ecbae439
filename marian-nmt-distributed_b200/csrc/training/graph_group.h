// Graph groups: one training update = forward, backward, gradient exchange,
// optimizer step.
//
// SingletonGraph follows the reference's src/training/graph_group_singleton.cu:21-66.
// SyncGraphGroup is the data-parallel group of src/training/graph_group_sync.cu:42-188
// re-cut for ONE PROCESS PER GPU: each rank owns one ExpressionGraph and the
// parameter shard `rank` (shardSize = total/N on the padded arena) with its own
// optimizer state.  An update is
//     computeGradients(batch_rank)             forward+backward (CUDA-graph replay)
//     reduce-scatter(sum) of the flat gradient arena      [exchange, see below]
//     updateShard()      fused {x 1/N, shard-norm clip, Adam} on the owned shard
//     all-gather of the flat parameter arena              [exchange]
// which is the reference's gather/add/update/scatter loop (:125-151) expressed
// as collectives.  The exchange itself is injected (ShardExchange): the product
// harness drives NCCL through torch.distributed on the engine stream, tests
// drive gloo over the CPU oracle - the group logic is the same code.
// Kept reference semantics: gradients are divided by N regardless of per-rank
// sentence counts; clipping uses the SHARD's norm; cost is the mean over ranks.
#pragma once

#include <chrono>
#include <map>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "data/batch.h"
#include "graph/expression_graph.h"
#include "models/model_factory.h"
#include "optimizers/optimizers.h"
#include "training/graph_replay.h"

namespace marian {

class GraphGroup {
protected:
  Ptr<Options> options_;
  Ptr<OptimizerBase> opt_;

public:
  GraphGroup(Ptr<Options> options) : options_(options), opt_(Optimizer(options)) {}
  virtual ~GraphGroup() {}
  virtual void update(Ptr<data::CorpusBatch>) = 0;
  virtual float cost() = 0;
};

// forward + backward of one batch on one graph, eagerly or by graph replay
class GradientWorker {
public:
  GradientWorker(Ptr<Options> options, int device) {
    graph_ = New<ExpressionGraph>();
    graph_->setDevice(device);
    graph_->reserveWorkspaceMB(options->get<size_t>("workspace"));
    builder_ = models::from_options(options);
    pinnedCost_ = (float*)device::mallocPinned(sizeof(float));
    *pinnedCost_ = 0.f;
    replay_.setEnabled(options->get<bool>("graph-replay", true));
  }
  ~GradientWorker() {
    replay_.clear();
    device::freePinned(pinnedCost_);
    device::freeMarker(eagerDone_);
  }

  Ptr<ExpressionGraph> graph() { return graph_; }
  Ptr<EncoderDecoder> builder() { return builder_; }
  StepReplay& replay() { return replay_; }
  // hook between the two halves of the backward sweep (SyncGraphGroup: first phase of the gradient exchange)
  void setMidStep(StepReplay::MidStep* mid) { mid_ = mid; }

  // Enqueues forward+backward on the engine stream; cost lands in pinned memory.
  void computeGradients(Ptr<data::CorpusBatch> batch, bool keepLogits = false) {
    device::setDevice((int)graph_->getDevice());
    gemmPrepareStep(graph_->getBackend()->getGemmHandle());  // BF16S mode: bf16 copy of the parameters up to date
    gemmAllowShadowOnly(graph_->getBackend()->getGemmHandle(), !keepLogits);  // a kept step leaves every fp32 tensor readable
    auto key = batch->shapeKey();
    if(!keepLogits) {
      if(auto plan = replay_.find(key)) {
        replay_.replay(*plan, *batch, mid_);
        lastReplayed_ = true;
        return;
      }
    }
    lastReplayed_ = false;
    // an eager build resets and refills the graph's own pinned staging: uploads of the previous eager step must
    // have executed (see StepReplay::replay)
    device::waitMarker(eagerDone_);
    bool capture = !keepLogits && replay_.shouldCapture(key);
    static const bool timing = std::getenv("MRN_NODE_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto costNode = builder_->build(graph_, batch);
    auto t1 = std::chrono::steady_clock::now();
    // split step: the group's hook runs between the two halves of the backward sweep - directly when the step is
    // eager, between the two graphs when it is captured (the capture is cut in two at the same point)
    void* execFirst = nullptr;
    size_t kernelsFirst = 0;
    if(mid_) {
      auto* mid = mid_;
      graph_->setBackwardSplit([mid](const std::list<Expr>& tape) { return mid->choose(tape); },
                               [mid, capture, &execFirst, &kernelsFirst]() {
                                 if(capture) {
                                   execFirst = device::endCapture();
                                   ABORT_IF(!execFirst, "CUDA graph capture of the first half of the training step failed");
                                   kernelsFirst = device::lastCaptureKernelCount();
                                   device::beginCapture();
                                 } else {
                                   mid->run();
                                 }
                               });
    } else {
      graph_->setBackwardSplit(nullptr, nullptr);
    }
    if(capture)
      device::beginCapture();
    graph_->forward();
    auto t2 = std::chrono::steady_clock::now();
    device::copyD2H(pinnedCost_, costNode->val()->data(), sizeof(float));
    if(keepLogits)
      logits_ = builder_->lastLogits();
    graph_->backward();
    if(timing) {
      auto t3 = std::chrono::steady_clock::now();
      auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
      fprintf(stderr, "[node-timing] build %.3f s, forward %.3f s, backward %.3f s (host wall time)\n", sec(t0, t1), sec(t1, t2), sec(t2, t3));
    }
    if(capture) {
      void* exec = device::endCapture();
      ABORT_IF(!exec, "CUDA graph capture of the training step failed");
      auto& plan = execFirst ? replay_.store(key, execFirst, graph_, exec, mid_->tag(), kernelsFirst) : replay_.store(key, exec, graph_);
      replay_.launch(plan, mid_);
    } else {
      eagerDone_ = device::recordMarker(eagerDone_);
    }
  }

  // Blocks until the enqueued work is done and returns the batch cost.
  float cost() {
    device::setDevice((int)graph_->getDevice());
    device::synchronize();
    return *pinnedCost_;
  }

  Expr logits() { return logits_; }
  bool lastStepReplayed() const { return lastReplayed_; }

private:
  Ptr<ExpressionGraph> graph_;
  Ptr<EncoderDecoder> builder_;
  StepReplay replay_;
  float* pinnedCost_;
  Expr logits_;
  bool lastReplayed_{false};
  void* eagerDone_{nullptr};  // marker behind the latest eager step (its uploads have read the graph's staging)
  StepReplay::MidStep* mid_{nullptr};
};

class SingletonGraph : public GraphGroup {
public:
  SingletonGraph(Ptr<Options> options, int device = 0) : GraphGroup(options), worker_(options, device) {}

  void update(Ptr<data::CorpusBatch> batch) {
    worker_.computeGradients(batch);
    opt_->update(worker_.graph());
  }
  float cost() { return worker_.cost(); }

  GradientWorker& worker() { return worker_; }
  Ptr<OptimizerBase> optimizer() { return opt_; }

private:
  GradientWorker worker_;
};

// Injected collective step of SyncGraphGroup (NCCL in the product harness, gloo
// in CPU tests, or in-process loops for single-process multi-rank simulation).
struct ShardExchange {
  virtual ~ShardExchange() {}
  // sum over ranks of flatGrads[rank*shard .. +shard) into shardOut on each rank
  virtual void reduceScatter(float* flatGrads, float* shardOut, size_t shardElements) = 0;
  // every rank's flatParams[rank*shard .. +shard) to all ranks, in place
  virtual void allGather(float* flatParams, size_t shardElements) = 0;
  // rank 0's buffer to all ranks
  virtual void broadcast(float* flat, size_t elements) = 0;
  virtual float meanCost(float localCost) = 0;
};

class SyncGraphGroup : public GraphGroup, public StepReplay::MidStep {
public:
  SyncGraphGroup(Ptr<Options> options, int device, int rank, int nranks, Ptr<ShardExchange> exchange = nullptr)
      : GraphGroup(options), worker_(options, device), rank_(rank), nranks_(nranks), exchange_(exchange) {
    worker_.graph()->params()->setShardCount(nranks);
    adam_ = std::dynamic_pointer_cast<Adam>(opt_);
    // piece-wise exchange overlapped with the backward sweep: ranks in powers of two (equal 16-byte aligned pieces)
    overlap_ = adam_ && (nranks == 2 || nranks == 4 || nranks == 8) && options->get<bool>("exchange-overlap", true) && std::getenv("MRN_NO_EXCHANGE_OVERLAP") == nullptr;
  }
  ~SyncGraphGroup() {
    worker_.setMidStep(nullptr);
    if(stampCount_ > 0 && rank_ == 0) {
      double n = (double)stampCount_;
      fprintf(stderr, "[exchange timing, rank 0, us, mean of %d sampled steps] early phase: barrier %.1f gather %.1f publish+barrier %.1f adam+stores %.1f | "
                      "late phase: barrier %.1f gather %.1f publish+barrier %.1f adam+stores %.1f | early start -> sweep end %.1f, early phase past sweep end %.1f, sweep end -> update end %.1f\n",
              stampCount_, stampSum_[0] / n, stampSum_[1] / n, stampSum_[2] / n, stampSum_[3] / n, stampSum_[4] / n, stampSum_[5] / n, stampSum_[6] / n, stampSum_[7] / n, stampSum_[8] / n,
              stampSum_[9] / n, stampSum_[10] / n);
    }
    if(partials_) {
      device::setDevice((int)worker_.graph()->getDevice());
      device::synchronize();
      device::freeDevice(partials_);
    }
  }

  void setExchange(Ptr<ShardExchange> e) { exchange_ = e; }

  // `batch` is this rank's part: batch->split(N)[rank] of the global batch
  // (reference :43), or a full per-rank batch for weak scaling.
  void update(Ptr<data::CorpusBatch> batch) {
    ABORT_IF(!exchange_, "SyncGraphGroup::update needs a ShardExchange");
    computeGradients(batch);
    if(first_) {
      // reference :46-53: all replicas start from graph 0's parameters before any gradient is
      // computed.  The pass above created and initialised the parameters; broadcast, then take
      // the first gradients again at the common parameter point.
      exchange_->broadcast(flatParams()->data(), flatParams()->size());
      first_ = false;
      computeGradients(batch);
    }
    exchange_->reduceScatter(flatGrads()->data(), shardGrads()->data(), shardSize());
    updateShard();
    exchange_->allGather(flatParams()->data(), shardSize());
  }

  // --- the phases, individually callable so a host harness can interleave its
  //     own collectives (bench.py / tests drive torch.distributed here) -------
  void computeGradients(Ptr<data::CorpusBatch> batch) {
    worker_.computeGradients(batch);
    ensureShard();
  }

  // fused {x 1/N, clip by shard norm, optimizer} on params[rank*shard ..) using
  // the summed gradient shard in shardGrads()
  void updateShard() {
    ensureShard();
    auto p = flatParams()->subtensor((int)(rank_ * shardSize_), (int)shardSize_);
    opt_->update(p, shardGrads_, 1.f, 1.f / (float)nranks_);
    gemmInvalidateCache(worker_.graph()->getBackend()->getGemmHandle());
    gemmParamsUpdated(worker_.graph()->getBackend()->getGemmHandle(), false);  // replicas / shards changed: bf16 arena copy is stale
  }

  // ---- peer-memory exchange (kernels/exchange.cu): reduce-scatter + Adam + all-gather as
  //      {barrier, gather-reduce by peer loads, Adam with peer stores, barrier} ----------------
  // signalPad(): 256 bytes of device memory other ranks write their barrier flags into
  void* signalPad() {
    if(!pad_) {
      device::setDevice((int)worker_.graph()->getDevice());
      pad_ = device::mallocDevice(kSignalPadBytes);
      device::zero(pad_, kSignalPadBytes);
      device::synchronize();
    }
    return pad_;
  }
  // tables: address of every rank's parameter arena / gradient arena / signal pad as mapped into
  // THIS process (own entries = local pointers)
  void setPeers(const PeerTable& params, const PeerTable& grads, const PeerTable& pads) {
    peerParams_ = params;
    peerGrads_ = grads;
    peerPads_ = pads;
    peersSet_ = true;
    if(overlap_) {
      device::setDevice((int)worker_.graph()->getDevice());
      partials_ = (float*)device::mallocDevice(256);
      device::zero(partials_, 256);
      worker_.setMidStep(this);  // steps built from now on are split where the upper shards' gradients are final
    }
  }
  bool overlapped() const { return overlap_ && peersSet_; }

  // ---- overlapped, piece-wise exchange --------------------------------------------------------------------
  // The arena keeps the reference's N contiguous shards as the units of the clipping norm (graph_group_sync.cu:
  // 130-142), but every rank owns PIECE `rank` of EVERY shard (and the Adam moments of those pieces).  The backward
  // sweep produces the gradients of the arena's tail first (decoder, output layer) and of its head last (encoder,
  // source embeddings): once every parameter at or above shard s0 is final the sweep is cut (StepReplay::MidStep),
  // and while its rest runs, the side stream exchanges shards [s0, N):
  //   barrier (all ranks are past the cut) -> gather-reduce of the own pieces by peer loads -> partial sums of squares
  //   to all ranks -> barrier -> clip by the shard norm + Adam + peer stores of the new parameters.
  // Shards [0, s0) follow after the sweep, then one barrier ends the update.
  size_t choose(const std::list<Expr>& tape) {
    split_ = -1;
    auto params = worker_.graph()->params();
    if(!overlapped() || params->size() == 0)
      return (size_t)-1;
    // last position in the sweep (0 = first node swept = last node of the tape) at which a parameter still receives gradient
    std::map<Chainable<Tensor>*, size_t> lastUse;
    size_t pos = 0;
    for(auto it = tape.rbegin(); it != tape.rend(); ++it, ++pos)
      if((*it)->trainable())
        for(auto& c : (*it)->children())
          if(c->type() == "param" && c->trainable())
            lastUse[c.get()] = pos;
    const size_t sweep = pos;
    const float* base = params->grads()->data();
    const size_t total = params->grads()->size(), shard = total / nranks_;
    // finalAt[s] = sweep position after which every parameter overlapping [s * shard, total) is final
    std::vector<size_t> finalAt(nranks_, 0);
    for(auto p : *params) {
      auto it = lastUse.find(p.get());
      if(it == lastUse.end() || !p->grad())
        continue;
      size_t end = (size_t)(p->grad()->data() - base) + p->grad()->size();
      for(int s = 0; s < nranks_; ++s)
        if(end > (size_t)s * shard)
          finalAt[s] = std::max(finalAt[s], it->second + 1);
    }
    // the lowest shard boundary whose tail is final with at least a fifth of the sweep still to run
    for(int s0 = 1; s0 < nranks_; ++s0)
      if(finalAt[s0] > 0 && finalAt[s0] * 5 <= sweep * 4) {
        split_ = s0;
        return finalAt[s0];
      }
    return (size_t)-1;
  }
  int tag() { return split_; }
  void restore(int tag) { split_ = tag; }
  void run() {  // between the two halves of the sweep: shards [split_, N) on the side stream
    if(split_ <= 0)
      return;
    device::forkSide();
    exchangePhase(0, split_, nranks_);
    device::returnFromSide();
    midRan_ = true;
  }

  bool peersSet() const { return peersSet_; }

  // after computeGradients(): everything else of the update, on the engine stream
  void exchangeUpdatePeer() {
    ABORT_IF(!peersSet_, "exchangeUpdatePeer: peers have not been mapped");
    ensureShard();
    if(overlap_) {
      stamp(10);
      device::joinSide();  // the first phase (if the step was split) is done before the second starts
      exchangePhase(1, 0, midRan_ ? split_ : nranks_);
      PeerBarrier(peerPads_, rank_, nranks_, ++epoch_);  // every piece has landed everywhere
      stamp(11);
      if(stamps_ && (++stampEvery_ % 8) == 0)
        collectStamps();
      midRan_ = false;
      stepBegun_ = false;
      gemmInvalidateCache(worker_.graph()->getBackend()->getGemmHandle());
      gemmParamsUpdated(worker_.graph()->getBackend()->getGemmHandle(), false);
      return;
    }
    int device = (int)worker_.graph()->getDevice();
    PeerBarrier(peerPads_, rank_, nranks_, ++epoch_);  // every rank finished backward
    PeerGatherReduce(shardGrads_, opt_->normSqScratch(device), peerGrads_, nranks_, (size_t)rank_ * shardSize_);
    auto p = flatParams()->subtensor((int)(rank_ * shardSize_), (int)shardSize_);
    PeerStores stores;
    stores.params = peerParams_;
    stores.nranks = nranks_;
    stores.self = rank_;
    stores.offset = (size_t)rank_ * shardSize_;
    opt_->updateShardWithPeers(p, shardGrads_, 1.f / (float)nranks_, stores);
    PeerBarrier(peerPads_, rank_, nranks_, ++epoch_);  // every shard has landed everywhere
    gemmInvalidateCache(worker_.graph()->getBackend()->getGemmHandle());
    gemmParamsUpdated(worker_.graph()->getBackend()->getGemmHandle(), false);  // replicas / shards changed: bf16 arena copy is stale
  }

  float cost() {
    float c = worker_.cost();
    return exchange_ ? exchange_->meanCost(c) : c;
  }
  float localCost() { return worker_.cost(); }

  Tensor flatParams() { return worker_.graph()->params()->vals(); }
  Tensor flatGrads() { return worker_.graph()->params()->grads(); }
  Tensor shardGrads() {
    ensureShard();
    return shardGrads_;
  }
  size_t shardSize() {
    ensureShard();
    return shardSize_;
  }
  int rank() const { return rank_; }
  int nranks() const { return nranks_; }
  GradientWorker& worker() { return worker_; }
  Ptr<OptimizerBase> optimizer() { return opt_; }

private:
  // one phase of the piece-wise exchange for reference shards [sa, sb), on the current stream
  void exchangePhase(int phase, int sa, int sb) {
    ensureShard();
    int device = (int)worker_.graph()->getDevice();
    if(!stepBegun_) {  // once per update: Adam step counter and bias corrections
      pieceArgs_ = adam_->beginPieceStep(shardSize_, device, 1.f / (float)nranks_);
      stepBegun_ = true;
    }
    PieceList pl;
    pl.len = shardSize_ / nranks_;
    for(int s = sa; s < sb; ++s) {
      pl.off[pl.count] = (size_t)s * shardSize_ + (size_t)rank_ * pl.len;
      pl.state[pl.count] = (size_t)s * pl.len;
      pl.shard[pl.count] = s;
      pl.count++;
    }
    stamp(phase * 5 + 0);
    PeerBarrier(peerPads_, rank_, nranks_, ++epoch_);  // every rank's gradients of these shards are final
    stamp(phase * 5 + 1);
    if(pl.count > 0) {
      device::zero(partials_ + phase * 8, 8 * sizeof(float));
      PeerGatherReducePieces(shardGrads_, partials_ + phase * 8, peerGrads_, nranks_, pl, phase == 0);
    }
    stamp(phase * 5 + 2);
    PeerPublishPartials(partials_ + phase * 8, peerPads_, rank_, nranks_, phase);
    PeerBarrier(peerPads_, rank_, nranks_, ++epoch_);  // all partial sums of squares are visible
    stamp(phase * 5 + 3);
    if(pl.count > 0)
      AdamUpdatePieces(peerParams_, pad_, rank_, nranks_, phase, shardGrads_, adam_->mt(), adam_->vt(), pieceArgs_, pl, phase == 0);
    stamp(phase * 5 + 4);
  }
  // tuning aid (MRN_EXCHANGE_TIMING=1): device time stamps around the pieces of the exchange; averages printed by rank 0
  // when the group is destroyed.  Slots: phase p -> 5p + {0 start, 1 after barrier, 2 after gather-reduce, 3 after publish +
  // barrier, 4 after Adam + peer stores}; 10 = second half of the sweep done (main stream), 11 = after the final barrier.
  void stamp(int slot) {
    static const bool on = std::getenv("MRN_EXCHANGE_TIMING") != nullptr;
    if(!on)
      return;
    if(!stamps_) {
      stamps_ = (unsigned long long*)device::mallocDevice(16 * 8);
      device::zero(stamps_, 16 * 8);
    }
    DeviceTimeStamp(stamps_ + slot);
  }
  void collectStamps() {
    if(!stamps_)
      return;
    unsigned long long h[16];
    device::synchronize();
    void* pin = device::pinnedScratch(sizeof(h));
    device::copyD2H(pin, stamps_, sizeof(h));
    device::synchronize();
    std::memcpy(h, pin, sizeof(h));
    if(h[10] && h[11] && h[5] && h[9]) {
      auto d = [&](int a, int b) { return h[a] && h[b] && h[b] >= h[a] ? (double)(h[b] - h[a]) / 1000.0 : 0.0; };
      double v[8] = {d(0, 1), d(1, 2), d(2, 3), d(3, 4), d(5, 6), d(6, 7), d(7, 8), d(8, 9)};
      for(int i = 0; i < 8; ++i)
        stampSum_[i] += v[i];
      stampSum_[8] += h[0] ? d(0, 10) : 0.0;   // start of the early phase .. end of the sweep
      stampSum_[9] += h[0] ? (h[4] > h[10] ? (double)(h[4] - h[10]) / 1000.0 : 0.0) : 0.0;  // early phase running past the sweep
      stampSum_[10] += d(10, 11);              // end of the sweep .. end of the update
      stampCount_++;
    }
  }

  void ensureShard() {
    if(shardGrads_)
      return;
    size_t total = flatParams()->size();
    ABORT_IF(total % nranks_ != 0, "parameter arena is not divisible into shards");
    shardSize_ = total / nranks_;
    shardAlloc_ = New<TensorAllocator>((int)worker_.graph()->getDevice());
    shardAlloc_->reserveExact(shardSize_ * sizeof(float));
    shardAlloc_->allocate(shardGrads_, Shape{1, (int)shardSize_});
    shardGrads_->set(0);
  }

  GradientWorker worker_;
  int rank_, nranks_;
  Ptr<ShardExchange> exchange_;
  bool first_{true};
  size_t shardSize_{0};
  Ptr<TensorAllocator> shardAlloc_;
  Tensor shardGrads_;
  // peer-memory exchange
  PeerTable peerParams_{}, peerGrads_{}, peerPads_{};
  bool peersSet_{false};
  int epoch_{0};
  void* pad_{nullptr};
  // overlapped piece-wise exchange
  Ptr<Adam> adam_;
  bool overlap_{false};
  int split_{-1};          // first reference shard of the early phase (-1: this step is not split)
  bool midRan_{false}, stepBegun_{false};
  float* partials_{nullptr};  // [phase 2][shard 8] sums of squares of the own pieces
  unsigned long long* stamps_{nullptr};
  double stampSum_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int stampCount_{0}, stampEvery_{0};
  AdamArgs pieceArgs_;
};

// Asynchronous SGD with a sharded parameter server - the reference's AsyncGraphGroup
// (src/training/graph_group_async.{h,cu}) re-cut for ONE PROCESS PER GPU over peer memory.
//
// Reference: a host thread per GPU; every update it (a) every tau steps fetches all parameter
// shards from their owner devices under a per-shard mutex, (b) runs forward/backward on its own
// batch, (c) pushes its gradient, shard by shard, to the owner device where that shard's optimizer
// updates the master copy under the same mutex.  No barrier between workers: stale gradients.
//
// Here rank r owns master shard r = one device allocation [lock, steps | p | m | v] mapped into
// every rank (CUDA IPC).  A worker step is a sequence of kernels on its own engine stream:
//   fetch  (every tau):  for each shard s: ShardLock(s) -> copy master_s.p into the local replica
//                        (peer read over NVLink) -> ShardUnlock(s)
//   computeGradients     CUDA-graph replay of forward + backward
//   push   (every tau):  for each shard s: sum of squares of the local slice, ShardLock(s, count
//                        the Adam step) -> AdamUpdateRemote (clip by the SLICE's norm + Adam on
//                        the owner's p, m, v through peer loads/stores) -> ShardUnlock(s)
// The spin lock lives in the owner's memory and is taken with system-scope atomics, so the
// pusher updates the master shard directly - the owner process is not involved.  Locks are
// taken one at a time in shard order: no hold-and-wait, no deadlock.  Adam only.
// Not carried over: exponential smoothing, learning-rate scaling by batch words (both off by
// default), scheduler quiescing for save/validate (control plane, out of scope).
class AsyncGraphGroup : public GraphGroup {
public:
  AsyncGraphGroup(Ptr<Options> options, int device, int rank, int nranks)
      : GraphGroup(options), worker_(options, device), rank_(rank), nranks_(nranks), tau_(std::max<size_t>(1, options->get<size_t>("optimizer-delay", 1))) {
    worker_.graph()->params()->setShardCount(nranks);
    adam_ = std::dynamic_pointer_cast<Adam>(opt_);
    ABORT_IF(!adam_, "AsyncGraphGroup: the remote shard update is implemented for adam");
  }
  // Peers push into / fetch from this rank's master block through their IPC mappings without any host
  // involvement of this process: the harness must stop all ranks (AsyncTrainer.close(): barrier) before one of them
  // is destroyed.  This rank's own queued kernels are drained here.
  ~AsyncGraphGroup() {
    if(master_) {
      device::setDevice((int)worker_.graph()->getDevice());
      device::synchronize();
      device::freeDevice(master_);
    }
  }

  // Builds the tape once (parameters initialised from the shared seed, identical on all ranks),
  // allocates this rank's master block and seeds it with its shard (reference init(): :96-147).
  void init(Ptr<data::CorpusBatch> batch) {
    if(master_)
      return;
    worker_.computeGradients(batch);  // eager first pass: parameters exist, gradients allocated
    device::setDevice((int)worker_.graph()->getDevice());
    size_t total = flatParams()->size();
    ABORT_IF(total % nranks_ != 0, "parameter arena is not divisible into shards");
    shardSize_ = total / nranks_;
    masterBytes_ = 256 + 3 * shardSize_ * sizeof(float);
    master_ = device::mallocDevice(masterBytes_);
    device::zero(master_, masterBytes_);
    device::copyD2D((uint8_t*)master_ + 256, flatParams()->data() + (size_t)rank_ * shardSize_, shardSize_ * sizeof(float));
    scratch_ = New<TensorAllocator>((int)worker_.graph()->getDevice());
    scratch_->reserveExact(512);
    scratch_->allocate(normSq_, Shape{1, 1});
    scratch_->allocate(steps_, Shape{1, 1});  // int32 step number of the shard being updated
    if(tau_ > 1) {
      accAlloc_ = New<TensorAllocator>((int)worker_.graph()->getDevice());
      accAlloc_->reserveExact(total * sizeof(float));
      accAlloc_->allocate(accGrads_, Shape{1, (int)total});
      accGrads_->set(0);
    }
    device::synchronize();
  }
  void* masterBlock() { return master_; }
  size_t masterBytes() const { return masterBytes_; }
  void setPeers(const PeerTable& masters) {
    masters_ = masters;
    peersSet_ = true;
  }

  void fetchParams() {
    ABORT_IF(!peersSet_, "AsyncGraphGroup: master shards have not been mapped");
    device::setDevice((int)worker_.graph()->getDevice());
    for(int s = 0; s < nranks_; ++s) {
      ShardLock(masters_.ptr[s], false, nullptr);
      device::copyD2D(flatParams()->data() + (size_t)s * shardSize_, (uint8_t*)masters_.ptr[s] + 256, shardSize_ * sizeof(float));
      ShardUnlock(masters_.ptr[s]);
    }
    gemmInvalidateCache(worker_.graph()->getBackend()->getGemmHandle());
    gemmParamsUpdated(worker_.graph()->getBackend()->getGemmHandle(), false);  // replicas / shards changed: bf16 arena copy is stale
  }

  void pushGradients(Tensor gradients) {
    ABORT_IF(!peersSet_, "AsyncGraphGroup: master shards have not been mapped");
    device::setDevice((int)worker_.graph()->getDevice());
    AdamArgs a = adam_->hyper();
    for(int s = 0; s < nranks_; ++s) {
      auto slice = gradients->subtensor((int)((size_t)s * shardSize_), (int)shardSize_);
      if(a.clipNorm > 0)
        SumSquares(normSq_, slice);  // the reference clips with the pushed shard's own norm
      ShardLock(masters_.ptr[s], true, (int*)steps_->data());
      AdamUpdateRemote(masters_.ptr[s], shardSize_, slice->data(), a, (const int*)steps_->data(), a.clipNorm > 0 ? normSq_ : nullptr);
      ShardUnlock(masters_.ptr[s]);
    }
  }

  // reference execute(): :150-215
  void update(Ptr<data::CorpusBatch> batch) {
    init(batch);
    if(t_ % tau_ == 0)
      fetchParams();
    worker_.computeGradients(batch);
    Tensor gradients = flatGrads();
    if(tau_ > 1) {
      using namespace functional;
      Element(_1 += _2, accGrads_, flatGrads());
      gradients = accGrads_;
    }
    t_++;
    if(t_ % tau_ == 0) {
      pushGradients(gradients);
      if(tau_ > 1)
        accGrads_->set(0);
    }
  }

  float cost() { return worker_.cost(); }
  Tensor flatParams() { return worker_.graph()->params()->vals(); }
  Tensor flatGrads() { return worker_.graph()->params()->grads(); }
  GradientWorker& worker() { return worker_; }
  size_t shardSize() const { return shardSize_; }
  int rank() const { return rank_; }
  int nranks() const { return nranks_; }

private:
  GradientWorker worker_;
  int rank_, nranks_;
  size_t tau_;
  size_t t_{0};
  Ptr<Adam> adam_;
  void* master_{nullptr};
  size_t masterBytes_{0};
  size_t shardSize_{0};
  PeerTable masters_{};
  bool peersSet_{false};
  Ptr<TensorAllocator> scratch_, accAlloc_;
  Tensor normSq_, steps_, accGrads_;
};

}  // namespace marian
