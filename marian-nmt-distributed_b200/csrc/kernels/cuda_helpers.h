// Small CUDA utilities shared by the kernel translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

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
// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// The training step is a chain of several hundred short kernels; between two of them the GPU
// idles for the launch latency of the second.  Kernels launched through launchPdl() may be
// SCHEDULED while their predecessor in the stream is still running: everything a kernel does
// before pdlWait() (index math, barrier/TMEM set-up, descriptor prefetch) overlaps the
// predecessor's tail; pdlWait() returns once the predecessor has completed and its writes are
// visible, so every global-memory access must come after it.  pdlTrigger() (issued right at the
// start) lets the NEXT kernel begin its own launch as soon as all CTAs of this one are resident.
// Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdlWait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdlTrigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdlEnter() {
  pdlTrigger();
  pdlWait();
}

// Measured on the Transformer-base step (B200, graph replay): 8.69 ms with the attribute on every
// hot kernel vs 8.46 ms without - the early-scheduled dependents take shared memory / TMEM from
// the running kernel's CTAs and the graph's programmatic edges are not free.  Hence OPT-IN
// (MRN_PDL=1) until the trigger points are tuned per kernel.
inline bool pdlEnabled() {
  static const bool on = std::getenv("MRN_PDL") != nullptr;
  return on;
}

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-serialization attribute.
// Only for kernels that call pdlWait()/pdlEnter() before touching global memory.
template <class... KArgs, class... Args>
inline void launchPdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdlEnabled() ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

// one 128-bit reduction into L2 instead of four scalar atomics (sm_90+); p must be 16-byte aligned
__device__ __forceinline__ void redAdd4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

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
