// bf16 shadow copies of GEMM operands (GemmMode::BF16S) - CUDA product only.
//
// The tensor-core products of the hot path are bound by operand bytes moving L2 -> shared memory
// (~64 B/clk/SM): fp32 operands (kind::tf32) move twice the bytes of bf16 operands and run the
// tensor pipe at half rate.  In BF16S mode every fp32 tensor that the graph knows to be a GEMM
// operand gets a bf16 copy ("shadow") in the same element order, written by the kernel that
// produces the fp32 tensor (layer-norm, attention, swish, their gradients, the GEMM epilogue, the
// cross-entropy gradient, Adam for the weights).  The GEMM reads the shadows through TMA in all four
// transpose cases (K-major and MN-major SWIZZLE_128B descriptors); anything that arrives without a
// shadow is converted by one flat pass (shadow::ensure in gemm.cu).  Storage, accumulation and all
// other operators stay fp32.
//
// Protocol (flags live in MemoryPiece, tensors/tensor.h, so reshape views share them):
//   graph     sets shadowWanted on a node's value when a product consumes it, and on the adjoint of a
//             product node when that node has exactly ONE consumer (one writer of the adjoint);
//   producer  calls shadow::produce(t): nullptr unless wanted and enabled, otherwise the buffer to
//             fill (same indexing as the fp32 tensor), marked valid;
//   GEMM      shadow::ensure(): valid shadow, or convert now.
#pragma once

#include <cuda_bf16.h>

#include "tensors/tensor.h"

namespace marian {
namespace shadow {

bool enabled();
void setEnabled(bool on);
// Buffer to fill for `t` (whole memory piece, element i of the piece at [i]) or nullptr.
__nv_bfloat16* produce(const Tensor& t);
// After produce(t) returned a buffer: may the producer skip the fp32 stores of `t` altogether?  (The graph marked
// the tensor shadowOnly: all its readers are products.)  Returns the fp32 destination to use - nullptr = skip.
float* fp32Target(const Tensor& t, const __nv_bfloat16* producedShadow);
// MRN_SHADOW_KEEP_FP32=1 (or setSkipFp32(false)) keeps every fp32 tensor complete (debugging, tensors fetched by the host)
void setSkipFp32(bool on);

#if defined(__CUDACC__)
// 4 consecutive elements (8-byte store); sh may be null
__device__ __forceinline__ void store4(__nv_bfloat16* sh, size_t i, float4 v) {
  if(sh) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(sh + i) = u;
  }
}
__device__ __forceinline__ void store1(__nv_bfloat16* sh, size_t i, float v) {
  if(sh)
    sh[i] = __float2bfloat16_rn(v);
}
#endif

}  // namespace shadow
}  // namespace marian
