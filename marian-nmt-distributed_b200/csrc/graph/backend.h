// Per-graph device backend.
//
// Role of the reference's Backend/BackendGPU (src/graph/backend.h:5-8,
// src/graph/backend_gpu.h:20-55: cuBLAS handle + cuRAND generator per graph).
// Here it owns the GEMM context (tcgen05 kernels + packed-operand scratch, see
// kernels/tensor_operators.h) and the counter that seeds dropout masks.
#pragma once

#include "common/definitions.h"
#include "kernels/tensor_operators.h"

namespace marian {

class Backend {
public:
  Backend(int deviceId, size_t seed) : device_(deviceId), seed_(seed) {
    device::setDevice(deviceId);
    gemm_ = createGemmContext(deviceId);
  }
  ~Backend() {
    destroyGemmContext(gemm_);
    if(dropEpoch_)
      device::freeDevice(dropEpoch_);
  }

  void setDevice(size_t) { device::setDevice(device_); }
  int getDevice() const { return device_; }

  GemmHandle getGemmHandle() { return gemm_; }
  // source-compatible spelling used by node operators ported from Marian
  GemmHandle getCublasHandle() { return gemm_; }

  uint64_t nextDropoutSeed() { return (uint64_t)seed_ * 0x9E3779B97F4A7C15ULL + (++dropCounter_); }

  // Device counter mixed into every dropout seed; bumped once per forward pass that draws masks
  // (the bump is part of a captured step, so replays get fresh masks).
  const uint64_t* dropoutEpoch() {
    if(!dropEpoch_) {
      device::setDevice(device_);
      dropEpoch_ = (uint64_t*)device::mallocDevice(256);
      device::zero(dropEpoch_, 256);
    }
    if(!epochBumped_) {
      DropoutEpochBump(dropEpoch_);
      epochBumped_ = true;
    }
    return dropEpoch_;
  }
  void newForwardPass() { epochBumped_ = false; }

private:
  int device_;
  size_t seed_;
  uint64_t dropCounter_{0};
  uint64_t* dropEpoch_{nullptr};
  bool epochBumped_{false};
  GemmHandle gemm_{nullptr};
};

typedef Backend BackendGPU;

}  // namespace marian
