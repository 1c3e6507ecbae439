// Pinned host staging for host->device uploads.
//
// The reference uploads every constant / index vector with a blocking
// cudaMemcpy from pageable memory, and CopyRows/PasteRows even cudaMalloc a
// temporary per call (src/tensors/tensor.cu:29-43,
// src/kernels/tensor_operators.cu:688-698).  Here uploads are staged in pinned
// chunks owned by the expression graph and copied asynchronously, so that
//  (a) nothing in a training step blocks the host, and
//  (b) a captured CUDA graph can re-read the same pinned addresses on replay
//      after the host refilled them with the next batch (training/graph_replay.h).
#pragma once

#include <cstring>
#include <functional>
#include <vector>

#include "common/definitions.h"
#include "tensors/device.h"

namespace marian {

namespace data {
class CorpusBatch;
}

class Staging {
public:
  ~Staging() {
    for(auto& c : chunks_)
      device::freePinned(c.base);
  }

  // Returns `bytes` of pinned memory valid until reset().
  void* take(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    if(chunks_.empty() || chunks_[cur_].used + bytes > chunks_[cur_].size) {
      // advance to an existing chunk that fits, else allocate a new one
      size_t i = chunks_.empty() ? 0 : cur_ + 1;
      for(; i < chunks_.size(); ++i)
        if(chunks_[i].used + bytes <= chunks_[i].size)
          break;
      if(i >= chunks_.size()) {
        size_t sz = std::max(bytes, kChunk);
        Chunk c;
        c.base = (uint8_t*)device::mallocPinned(sz);
        c.size = sz;
        c.used = 0;
        chunks_.push_back(c);
        i = chunks_.size() - 1;
      }
      cur_ = i;
    }
    void* p = chunks_[cur_].base + chunks_[cur_].used;
    chunks_[cur_].used += bytes;
    return p;
  }

  void reset() {
    for(auto& c : chunks_)
      c.used = 0;
    cur_ = 0;
  }

  size_t bytesInUse() const {
    size_t s = 0;
    for(auto& c : chunks_)
      s += c.used;
    return s;
  }

private:
  struct Chunk {
    uint8_t* base;
    size_t size;
    size_t used;
  };
  static constexpr size_t kChunk = 4u << 20;
  std::vector<Chunk> chunks_;
  size_t cur_{0};
};

// A host->device upload whose content depends on the current batch.  `refill`
// rewrites the pinned source for a new batch of the same shape; the device
// copy itself is part of the recorded work.
struct BatchUpload {
  void* pinned;
  size_t bytes;
  std::function<void(void* pinned, const data::CorpusBatch&)> refill;
};

}  // namespace marian
