// Small CUDA utilities shared by the kernel translation units.
#pragma once

#include <cuda_runtime.h>

#include "common/definitions.h"
#include "tensors/device.h"

namespace marian {

// The reference aborts on CUDA errors (src/kernels/cuda_helpers.h:7-18); here
// they surface as MarianError and reach the caller of the C ABI as a status.
#define CUDA_CHECK(expr)                                                                     \
  do {                                                                                       \
    cudaError_t rc__ = (expr);                                                               \
    if(rc__ != cudaSuccess)                                                                  \
      ::marian::abort_with(__FILE__, __LINE__, std::string("CUDA error: ") + cudaGetErrorString(rc__)); \
  } while(0)

#define CUDA_LAUNCH_CHECK() CUDA_CHECK(cudaGetLastError())

inline cudaStream_t cudaStreamOfEngine() {
  return (cudaStream_t)device::currentStream();
}

// B200: 148 SMs.  Grid-stride kernels are sized in multiples of this.
constexpr int kNumSMs = 148;

inline int gridFor(size_t items, int threads, int blocksPerSM = 8) {
  size_t blocks = (items + threads - 1) / threads;
  size_t cap = (size_t)kNumSMs * blocksPerSM;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

#if defined(__CUDACC__)
__device__ __forceinline__ float warpSum(float v) {
#pragma unroll
  for(int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warpMax(float v) {
#pragma unroll
  for(int o = 16; o > 0; o >>= 1)
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `smem` holds 32 floats.
__device__ __forceinline__ float blockSum(float v, float* smem) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warpSum(v);
  __syncthreads();  // protect smem reuse across consecutive calls
  if(lane == 0)
    smem[warp] = v;
  __syncthreads();
  int nwarps = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nwarps) ? smem[threadIdx.x] : 0.f;
  if(warp == 0)
    r = warpSum(r);
  if(threadIdx.x == 0)
    smem[0] = r;
  __syncthreads();
  return smem[0];
}
__device__ __forceinline__ float blockMax(float v, float* smem) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warpMax(v);
  __syncthreads();
  if(lane == 0)
    smem[warp] = v;
  __syncthreads();
  int nwarps = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nwarps) ? smem[threadIdx.x] : -3.4e38f;
  if(warp == 0)
    r = warpMax(r);
  if(threadIdx.x == 0)
    smem[0] = r;
  __syncthreads();
  return smem[0];
}
#endif

}  // namespace marian
