// TensorBase: a float view (pointer + Shape + device id) into arena memory.
//
// Same surface as the reference's TensorBase (src/tensors/tensor.h:15-65:
// data/shape/size/memory/getDevice/subtensor/get/set/copyFrom/scalar) so that
// node operators and graph groups read the same.  Differences by design:
// every device operation is asynchronous on the engine stream; only the
// host-returning accessors (get/scalar) synchronise.
#pragma once

#include <cstring>
#include <vector>

#include "common/definitions.h"
#include "common/shape.h"
#include "tensors/device.h"
#include "tensors/staging.h"

namespace marian {

// A span of arena memory.  The workspace never relocates (chunked arena, see
// tensors/allocator.h), so unlike the reference (src/tensors/memory_piece.h)
// the pointer is immutable after construction.
class MemoryPiece {
public:
  MemoryPiece(uint8_t* data, size_t size) : data_(data), size_(size) {}
  uint8_t* data() const { return data_; }
  size_t size() const { return size_; }
  void set(uint8_t* data, size_t size) {
    data_ = data;
    size_ = size;
  }
  // see TensorBase::data(): shared by every tensor view over this piece (reshape nodes)
  mutable bool lazyZero{false};
  // bf16 shadow copy for the tensor-core GEMM (GemmMode::BF16S, kernels/shadow.h).  The GRAPH sets
  // shadowWanted on values / adjoints that will be GEMM operands (and, for adjoints, have exactly one
  // writer); the operator that writes the fp32 tensor then also writes the bf16 copy (`shadow`,
  // same element order) and sets shadowValid; the GEMM converts itself when nobody did.
  mutable void* shadow{nullptr};
  mutable bool shadowWanted{false};
  mutable bool shadowValid{false};
  mutable uint32_t shadowGen{0};  // pool generation the buffer belongs to (stale buffers are ignored)
  // lane mark behind a shadow that a CONSUMER converted (device::laneMark): a consumer on another lane waits for it
  // (shadows written by the producing kernel need none: consumers wait for the producing node anyway)
  mutable void* shadowMark{nullptr};
  // shadowOnly: every reader of this tensor is a tensor-core product (graph analysis, see Node): a producer that
  // writes the bf16 shadow may leave the fp32 bytes unwritten (half of the step's activation write traffic goes
  // to fp32 tensors nobody reads).  fp32Skipped records that one did - fp32 readers must not appear after that.
  mutable bool shadowOnly{false};
  mutable bool fp32Skipped{false};
  // arena generation this piece was cut from (Allocator::clear() starts a new one): a piece that outlives a clear() -
  // an Expr kept across graph->clear() - must not free the new owner of its old address
  uint32_t arenaEpoch{0};

private:
  uint8_t* data_;
  size_t size_;
};

// Pinned staging used by TensorBase::set(vector) on the calling thread.  An
// ExpressionGraph installs its own staging while it builds/runs.
Staging*& currentStagingSlot();
inline Staging* currentStaging() { return currentStagingSlot(); }

class TensorBase : public std::enable_shared_from_this<TensorBase> {
public:
  TensorBase(Ptr<MemoryPiece> memory, Shape shape, int deviceId)
      : memory_(memory), shape_(shape), device_(deviceId) {}

  // Lazy zero ("first writer assigns"): an adjoint starts as "logically all zeros, not yet
  // written".  The reference memsets every adjoint and then accumulates into it
  // (src/graph/node.cu:31-36 + a stream sync per set(0)); here an accumulating operator that
  // finds the flag set ASSIGNS instead (takeLazyZero) - no memset, no read of the old value.
  // Any other access through data() materialises the zeros first, so unaware code is safe.
  float* data() const {
    if(memory_->lazyZero) {
      memory_->lazyZero = false;
      device::setDevice(device_);
      device::zero(memory_->data(), (size_t)shape_.elements() * sizeof(float));
    }
    return (float*)memory_->data();
  }
  // the bytes as they are: no lazy-zero materialisation (for code that is about to overwrite them)
  float* rawData() const { return (float*)memory_->data(); }
  void setLazyZero() {
#ifdef MRN_ORACLE_CPU
    // the CPU oracle keeps the reference's memset-then-accumulate (its operators also call
    // data() from inside OpenMP regions, where a lazy memset would race)
    set(0.f);
#else
    memory_->lazyZero = true;
#endif
  }
  bool isLazyZero() const { return memory_->lazyZero; }
  // true: the tensor holds no defined values and the caller promises to overwrite ALL of it
  bool takeLazyZero() {
    bool f = memory_->lazyZero;
    memory_->lazyZero = false;
    return f;
  }
  const Shape& shape() const { return shape_; }
  size_t size() const { return (size_t)shape_.elements(); }
  Ptr<MemoryPiece> memory() const { return memory_; }
  int getDevice() const { return device_; }

  void reset(Ptr<MemoryPiece> memory) {
    memory_ = memory;
  }

  Tensor subtensor(int offset, int size) {
    auto mem = New<MemoryPiece>((uint8_t*)data() + sizeof(float) * (size_t)offset, sizeof(float) * (size_t)size);
    return Tensor(new TensorBase(mem, Shape{1, size}, device_));
  }

  // --- host-returning accessors: synchronise (not allowed during capture) ---
  float get(size_t i) {
    std::vector<float> one(1);
    fetch(data() + i, one.data(), 1);
    return one[0];
  }
  float scalar() {
    ABORT_IF(size() != 1, "Tensor is not a scalar");
    return get(0);
  }
  void get(std::vector<float>& v) {
    v.resize(size());
    fetch(data(), v.data(), size());
  }

  // --- asynchronous mutators ---
  void set(float value) {
    device::setDevice(device_);
    memory_->lazyZero = false;
    if(value == 0.f)
      device::zero(data(), size() * sizeof(float));
    else
      device::fill(data(), value, size());
  }
  void set(const float* begin, size_t n) { upload(begin, n * sizeof(float)); }
  void set(const std::vector<float>& v) {
    ABORT_IF(v.size() != size(), "set(vector): size mismatch", v.size(), size());
    upload(v.data(), v.size() * sizeof(float));
  }
  void set(size_t i, float value) {
    std::vector<float> one(1, value);
    device::setDevice(device_);
    Staging* st = currentStaging();
    ABORT_IF(!st, "set(i, value) requires an active staging scope");
    void* p = st->take(sizeof(float));
    std::memcpy(p, one.data(), sizeof(float));
    device::copyH2D(data() + i, p, sizeof(float));
  }
  void copyFrom(Tensor in) {
    ABORT_IF(in->size() != size(), "copyFrom: size mismatch");
    device::setDevice(device_);
    memory_->lazyZero = false;
    device::copyD2D(data(), in->data(), size() * sizeof(float));
  }

  // Uploads raw bytes through pinned staging; returns the pinned source so the
  // caller can register a BatchUpload for it.
  void* upload(const void* src, size_t bytes) {
    device::setDevice(device_);
    if(bytes >= size() * sizeof(float))
      memory_->lazyZero = false;
    Staging* st = currentStaging();
    // large one-off uploads (parameter initialisation) do not go through the
    // graph's pinned staging: they would pin hundreds of MB for nothing
    if(bytes > (1u << 20) && !device::capturing()) {
      device::copyH2DBlocking(data(), src, bytes);
      return nullptr;
    }
    if(st) {
      void* p = st->take(bytes);
      std::memcpy(p, src, bytes);
      device::copyH2D(data(), p, bytes);
      return p;
    }
    // no graph scope (tests, one-off initialisation): blocking upload
    device::copyH2DBlocking(data(), src, bytes);
    return nullptr;
  }

  std::string debug();

private:
  void fetch(const float* src, float* dst, size_t n) {
    ABORT_IF(device::capturing(), "Host read-back of a tensor during CUDA graph capture");
    device::setDevice(device_);
    void* p = device::pinnedScratch(n * sizeof(float));
    device::copyD2H(p, src, n * sizeof(float));
    device::synchronize();
    std::memcpy(dst, p, n * sizeof(float));
  }

  Ptr<MemoryPiece> memory_;
  Shape shape_;
  int device_;
};

inline Tensor operator<<(Tensor t, const std::vector<float>& v) {
  t->set(v);
  return t;
}
inline Tensor operator>>(Tensor t, std::vector<float>& v) {
  t->get(v);
  return t;
}

}  // namespace marian
