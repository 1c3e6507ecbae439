// Arena allocator for graph workspace and parameter storage.
//
// Role of the reference's Allocator<DeviceGPU> + TensorAllocator
// (src/tensors/allocator.h:69-226, src/tensors/tensor_allocator.h:13-78):
// 256-byte aligned sub-allocation from device arenas, alloc/free per node,
// AllocationException for --mini-batch-fit probing, asTensor() exposing a whole
// arena as ONE flat [1,n] tensor (parameters / gradients).
//
// B200-first differences: the arena is a list of CHUNKS that never move (the
// reference re-allocates and copies through the host on growth,
// src/tensors/device_gpu.cu:17-36).  Stable addresses are what lets a captured
// CUDA graph of a whole step be replayed; with 180 GB of HBM there is no
// reason to compact.  Free gaps are indexed by size (best fit) and by address
// (O(log n) coalescing) instead of the reference's linear scan.
#pragma once

#include <map>
#include <set>
#include <unordered_map>
#include <vector>

#include "common/definitions.h"
#include "tensors/device.h"
#include "tensors/tensor.h"

namespace marian {

class AllocationException : public std::exception {
public:
  virtual const char* what() const throw() { return "Memory re-allocation attempted"; }
};

class Allocator {
public:
  Allocator(int deviceId, size_t bytes, size_t step, size_t alignment = 256)
      : device_(deviceId), step_(step), alignment_(alignment) {
    if(bytes)
      reserve(bytes);
  }
  ~Allocator() { release(); }

  Allocator(const Allocator&) = delete;

  size_t align(size_t size) const { return (size + alignment_ - 1) / alignment_ * alignment_; }

  void throwAtReallocation(bool t) { throw_ = t; }

  // Drops everything and makes the arena one chunk of `bytes`.
  void reserve(size_t bytes) {
    bytes = align(bytes);
    if(chunks_.size() == 1 && chunks_[0].size == bytes) {
      clear();
      return;
    }
    release();
    addChunk(bytes);
    clear();
  }

  template <typename T>
  size_t capacity(size_t num) const {
    return align(num * sizeof(T));
  }

  template <typename T>
  Ptr<MemoryPiece> alloc(size_t num) {
    return alloc(capacity<T>(num));
  }

  Ptr<MemoryPiece> alloc(size_t bytes) {
    bytes = align(std::max(bytes, (size_t)1));
    auto it = bySize_.lower_bound({bytes, nullptr});
    if(it == bySize_.end()) {
      if(throw_)
        throw AllocationException();
      size_t add = std::max(step_, bytes);
      uint8_t* base = addChunk(add);
      insertGap(base, add);
      ++generation_;
      it = bySize_.lower_bound({bytes, nullptr});
    }
    size_t gsize = it->first;
    uint8_t* gptr = it->second;
    bySize_.erase(it);
    byAddr_.erase(gptr);
    if(gsize > bytes)
      insertGapRaw(gptr + bytes, gsize - bytes);
    allocated_[gptr] = bytes;
    inUse_ += bytes;
    peak_ = std::max(peak_, inUse_);
    auto piece = New<MemoryPiece>(gptr, bytes);
    piece->arenaEpoch = epoch_;
    return piece;
  }

  // While side-stream work may still read freed tensors (see ExpressionGraph::backward) frees
  // are parked; flushDeferred() performs them.
  void deferFrees(bool on) { defer_ = on; }
  void flushDeferred() {
    std::vector<Ptr<MemoryPiece>> parked;
    parked.swap(deferred_);
    bool was = defer_;
    defer_ = false;
    for(auto& mp : parked)
      free(mp);
    defer_ = was;
  }

  bool free(Ptr<MemoryPiece> mp) {
    if(!mp || !mp->data())
      return false;
    if(mp->arenaEpoch != epoch_) {  // allocated before the last clear(): its bytes already went back to the arena
      mp->set(nullptr, 0);
      return false;
    }
    if(defer_) {
      deferred_.push_back(mp);
      return true;
    }
    auto it = allocated_.find(mp->data());
    if(it == allocated_.end())
      return false;
    size_t bytes = it->second;
    allocated_.erase(it);
    inUse_ -= bytes;
    insertGap(mp->data(), bytes);
    mp->set(nullptr, 0);
    return true;
  }

  void clear() {
    ++epoch_;
    deferred_.clear();
    bySize_.clear();
    byAddr_.clear();
    allocated_.clear();
    inUse_ = 0;
    for(auto& c : chunks_)
      insertGapRaw(c.base, c.size);
  }

  // Whole arena as one piece; only meaningful for single-chunk arenas
  // (parameter/gradient storage reserved with the exact size).
  Ptr<MemoryPiece> memory() {
    ABORT_IF(chunks_.size() != 1, "memory(): arena is not a single chunk");
    return New<MemoryPiece>(chunks_[0].base, chunks_[0].size);
  }

  size_t size() const {
    size_t s = 0;
    for(auto& c : chunks_)
      s += c.size;
    return s;
  }
  size_t inUse() const { return inUse_; }
  size_t peak() const { return peak_; }
  int getDevice() const { return device_; }
  // Bumped whenever a chunk is added: captured graphs stay valid (addresses
  // never move) but callers may want to know the arena grew.
  size_t generation() const { return generation_; }

private:
  struct Chunk {
    uint8_t* base;
    size_t size;
  };

  uint8_t* addChunk(size_t bytes) {
    device::setDevice(device_);
    uint8_t* p = (uint8_t*)device::mallocDevice(bytes);
    chunks_.push_back({p, bytes});
    return p;
  }

  void release() {
    device::setDevice(device_);
    for(auto& c : chunks_)
      device::freeDevice(c.base);
    chunks_.clear();
    bySize_.clear();
    byAddr_.clear();
    allocated_.clear();
    inUse_ = 0;
  }

  // Both addresses are gap/allocation START addresses, hence strictly inside a chunk.
  bool sameChunk(uint8_t* a, uint8_t* b) const {
    for(auto& c : chunks_)
      if(a >= c.base && a < c.base + c.size)
        return b >= c.base && b < c.base + c.size;
    return false;
  }

  void insertGapRaw(uint8_t* ptr, size_t size) {
    byAddr_[ptr] = size;
    bySize_.insert({size, ptr});
  }

  // Insert with coalescing of address-adjacent gaps (never across chunks).
  void insertGap(uint8_t* ptr, size_t size) {
    auto next = byAddr_.lower_bound(ptr);
    if(next != byAddr_.end() && ptr + size == next->first && sameChunk(ptr, next->first)) {
      size += next->second;
      bySize_.erase({next->second, next->first});
      next = byAddr_.erase(next);
    }
    if(next != byAddr_.begin()) {
      auto prev = std::prev(next);
      if(prev->first + prev->second == ptr && sameChunk(prev->first, ptr)) {
        ptr = prev->first;
        size += prev->second;
        bySize_.erase({prev->second, prev->first});
        byAddr_.erase(prev);
      }
    }
    insertGapRaw(ptr, size);
  }

  int device_;
  size_t step_;
  size_t alignment_;
  bool throw_{false};
  std::vector<Chunk> chunks_;
  std::set<std::pair<size_t, uint8_t*>> bySize_;
  std::map<uint8_t*, size_t> byAddr_;
  std::unordered_map<uint8_t*, size_t> allocated_;
  size_t inUse_{0};
  uint32_t epoch_{0};
  size_t peak_{0};
  size_t generation_{0};
  bool defer_{false};
  std::vector<Ptr<MemoryPiece>> deferred_;
};

class TensorAllocator {
public:
  explicit TensorAllocator(int deviceId) : allocator_(New<Allocator>(deviceId, 0, GROW, ALIGN)) {}

  void throwAtReallocation(bool t) { allocator_->throwAtReallocation(t); }

  // Reference semantics (tensor_allocator.h:33-41): round up to the next
  // multiple of the 512 MB growth step.
  void reserve(size_t bytes = 0) {
    size_t mult = bytes / GROW + 1;
    allocator_->reserve(mult * GROW);
  }
  void reserveExact(size_t bytes = 0) { allocator_->reserve(bytes); }

  void clear() { allocator_->clear(); }

  size_t capacity(const Shape& shape) { return allocator_->capacity<float>(shape.elements()); }

  void allocate(Tensor& t, const Shape& shape) {
    if(!t || t->shape() != shape) {
      auto mem = allocator_->alloc<float>(shape.elements());
      t = Tensor(new TensorBase(mem, shape, allocator_->getDevice()));
    }
  }

  void free(Tensor& t) { allocator_->free(t->memory()); }

  Tensor asTensor() {
    auto mem = allocator_->memory();
    int size = (int)(mem->size() / sizeof(float));
    return Tensor(new TensorBase(mem, Shape{1, size}, allocator_->getDevice()));
  }

  size_t size() { return allocator_->size() / sizeof(float); }
  Ptr<Allocator> allocator() { return allocator_; }

private:
  static constexpr size_t GROW = 512u << 20;
  static constexpr size_t ALIGN = 256;
  Ptr<Allocator> allocator_;
};

}  // namespace marian
