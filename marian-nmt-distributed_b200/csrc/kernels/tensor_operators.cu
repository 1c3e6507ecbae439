// Hand-written sm_100a kernels for the row-wise, gather/scatter and recurrent
// tensor operators of the hot path.  Semantics follow the reference's
// src/kernels/tensor_operators.cu (cited per operator) including its quirks
// (assigning backward of TransposeND/Shift/Highway, LN-grad clip at +-1000 and
// x_hat recovered from y, float labels in cross-entropy).
//
// All of these are HBM-bound.  Common design:
//  * rows are processed by a warp (short rows) or a block (long rows) with
//    shuffle reductions - the reference runs a shared-memory tree with a
//    __syncthreads per level and re-reads each row 3 times from global memory;
//  * cross-entropy / softmax statistics are computed in ONE pass over the row
//    (online max/sum), the 128 KB logits row of the V=32000 case is read once
//    in forward and twice in backward (2nd read served by L2);
//  * column reductions (bias / gamma / beta / va gradients) accumulate in
//    registers per thread across the rows a block owns and issue ONE atomicAdd
//    per block and column - the reference issues one atomicAdd per element
//    (3200-way contention on 512 addresses for layer-norm);
//  * 128-bit accesses wherever rows are 16-byte aligned;
//  * grids are sized in multiples of 148 SMs; everything is launched on the
//    engine stream and nothing synchronises.
#include <cuda_runtime.h>

#include <algorithm>

#include "kernels/cuda_helpers.h"
#include "kernels/shadow.h"
#include "kernels/tensor_operators.h"

namespace marian {

namespace {

__device__ __forceinline__ float stableLogit(float x) {
  // reference: tensor_operators.cu:15-23
  if(x >= 0.f) {
    float z = expf(-x);
    return 1.0f / (1.0f + z);
  } else {
    float z = expf(x);
    return z / (1.0f + z);
  }
}

// ---- row-processing skeleton ------------------------------------------------
// WARP=true : one warp per row (blockDim = 256 -> 8 rows per block)
// WARP=false: one block per row, grid-stride over rows
template <bool WARP>
struct RowCtx {
  __device__ __forceinline__ static int firstRow() { return WARP ? blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : blockIdx.x; }
  __device__ __forceinline__ static int rowStride() { return WARP ? gridDim.x * (blockDim.x >> 5) : gridDim.x; }
  __device__ __forceinline__ static int firstCol() { return WARP ? (threadIdx.x & 31) : threadIdx.x; }
  __device__ __forceinline__ static int colStride() { return WARP ? 32 : blockDim.x; }
  __device__ __forceinline__ static float sum(float v, float* smem) { return WARP ? warpSum(v) : blockSum(v, smem); }
  __device__ __forceinline__ static float max(float v, float* smem) { return WARP ? warpMax(v) : blockMax(v, smem); }
  __device__ __forceinline__ static bool leader() { return WARP ? ((threadIdx.x & 31) == 0) : (threadIdx.x == 0); }
};

struct RowLaunch {
  bool warp;
  int grid;
  int block;
};
// rows shorter than this go one-warp-per-row
constexpr int kWarpRowMaxCols = 256;
inline RowLaunch rowLaunch(int rows, int cols) {
  RowLaunch l;
  l.warp = cols <= kWarpRowMaxCols;
  if(l.warp) {
    l.block = 256;
    int blocks = (rows + 7) / 8;
    l.grid = std::max(1, std::min(blocks, kNumSMs * 16));
  } else {
    l.block = cols >= 4096 ? 1024 : (cols >= 1024 ? 512 : 256);
    l.grid = std::max(1, std::min(rows, kNumSMs * (2048 / l.block)));
  }
  return l;
}

#define ROW_DISPATCH(KERNEL, L, ...)                                             \
  do {                                                                           \
    if((L).warp)                                                                 \
      KERNEL<true><<<(L).grid, (L).block, 0, cudaStreamOfEngine()>>>(__VA_ARGS__);  \
    else                                                                         \
      KERNEL<false><<<(L).grid, (L).block, 0, cudaStreamOfEngine()>>>(__VA_ARGS__); \
    CUDA_LAUNCH_CHECK();                                                         \
  } while(0)

// online (single pass) max / sum-of-exp update
__device__ __forceinline__ void onlineUpdate(float& m, float& s, float x) {
  if(x > m) {
    s = s * expf(m - x) + 1.f;
    m = x;
  } else {
    s += expf(x - m);
  }
}

struct MaskView {
  const float* p;
  int rb[3];
  int cs;
  int d1, d2;
};

}  // namespace

bool IsNan(Tensor) {
  return false;  // reference: stubbed to false (tensor_operators.cu:25-33)
}

// =============================================================================
// Softmax family          reference: tensor_operators.cu:202-519
// =============================================================================
namespace {

template <bool WARP>
__global__ void gSoftmax(float* __restrict__ out, const float* __restrict__ in, MaskView mask, int rows, int cols) {
  __shared__ float smem[32];
  typedef RowCtx<WARP> R;
  for(int j = R::firstRow(); j < rows; j += R::rowStride()) {
    float* so = out + (size_t)j * cols;
    const float* sp = in + (size_t)j * cols;
    const float* mrow = nullptr;
    if(mask.p) {
      int o2 = j % mask.d2;
      int t = j / mask.d2;
      int o1 = t % mask.d1;
      int o0 = t / mask.d1;
      mrow = mask.p + (size_t)o0 * mask.rb[0] + (size_t)o1 * mask.rb[1] + (size_t)o2 * mask.rb[2];
    }
    float m = -1.70141e+38f;  // the reference's CUDA_FLT_MAX sentinel
    for(int id = R::firstCol(); id < cols; id += R::colStride()) {
      float mv = mrow ? mrow[(size_t)id * mask.cs] : 1.f;
      float x = sp[id];
      if(mv && x > m)
        m = x;
    }
    m = R::max(m, smem);
    float s = 0.f;
    for(int id = R::firstCol(); id < cols; id += R::colStride()) {
      float mv = mrow ? mrow[(size_t)id * mask.cs] : 1.f;
      float ex = mv ? expf(sp[id] - m) : 0.f;
      so[id] = ex;
      s += ex;
    }
    s = R::sum(s, smem);
    for(int id = R::firstCol(); id < cols; id += R::colStride())
      so[id] = so[id] / s;
  }
}

template <bool WARP>
__global__ void gLogSoftmax(float* __restrict__ out, const float* __restrict__ in, int rows, int cols) {
  __shared__ float smem[32];
  typedef RowCtx<WARP> R;
  for(int j = R::firstRow(); j < rows; j += R::rowStride()) {
    float* so = out + (size_t)j * cols;
    const float* sp = in + (size_t)j * cols;
    float m = -3.4e38f, s = 0.f;
    for(int id = R::firstCol(); id < cols; id += R::colStride())
      onlineUpdate(m, s, sp[id]);
    float M = R::max(m, smem);
    s = R::sum(s * expf(m - M), smem);
    float lse = logf(s);
    for(int id = R::firstCol(); id < cols; id += R::colStride())
      so[id] = (sp[id] - M) - lse;
  }
}

template <bool WARP>
__global__ void gSoftmaxGrad(float* __restrict__ grad, const float* __restrict__ adj, const float* __restrict__ val, int rows, int cols) {
  __shared__ float smem[32];
  typedef RowCtx<WARP> R;
  for(int j = R::firstRow(); j < rows; j += R::rowStride()) {
    float* g = grad + (size_t)j * cols;
    const float* a = adj + (size_t)j * cols;
    const float* v = val + (size_t)j * cols;
    float s = 0.f;
    for(int id = R::firstCol(); id < cols; id += R::colStride())
      s += v[id] * a[id];
    s = R::sum(s, smem);
    for(int id = R::firstCol(); id < cols; id += R::colStride()) {
      float x = v[id] * (a[id] - s);
      if(x)
        g[id] += x;
    }
  }
}

template <bool WARP>
__global__ void gLogSoftmaxGrad(float* __restrict__ grad, const float* __restrict__ adj, const float* __restrict__ val, int rows, int cols) {
  __shared__ float smem[32];
  typedef RowCtx<WARP> R;
  for(int j = R::firstRow(); j < rows; j += R::rowStride()) {
    float* g = grad + (size_t)j * cols;
    const float* a = adj + (size_t)j * cols;
    const float* v = val + (size_t)j * cols;
    float s = 0.f;
    for(int id = R::firstCol(); id < cols; id += R::colStride())
      s += a[id];
    s = R::sum(s, smem);
    for(int id = R::firstCol(); id < cols; id += R::colStride())
      g[id] += a[id] - (expf(v[id]) * s);
  }
}

}  // namespace

void Softmax(Tensor out, Tensor in, Tensor mask) {
  device::setDevice(out->getDevice());
  int cols = out->shape().back();
  int rows = out->shape().elements() / cols;
  MaskView mv;
  mv.p = nullptr;
  mv.d1 = mv.d2 = 1;
  mv.cs = 0;
  mv.rb[0] = mv.rb[1] = mv.rb[2] = 0;
  if(mask) {
    Shape4 os(out->shape()), ms(mask->shape());
    mv.p = mask->data();
    mv.rb[0] = ms.bst[0];
    mv.rb[1] = ms.bst[1];
    mv.rb[2] = ms.bst[2];
    mv.cs = ms.bst[3];
    mv.d1 = os.d[1];
    mv.d2 = os.d[2];
  }
  auto l = rowLaunch(rows, cols);
  ROW_DISPATCH(gSoftmax, l, out->data(), in->data(), mv, rows, cols);
}

void LogSoftmax(Tensor out, Tensor in) {
  device::setDevice(out->getDevice());
  int cols = out->shape().back();
  int rows = out->shape().elements() / cols;
  auto l = rowLaunch(rows, cols);
  ROW_DISPATCH(gLogSoftmax, l, out->data(), in->data(), rows, cols);
}

void SoftmaxGrad(Tensor grad, Tensor adj, Tensor val) {
  device::setDevice(adj->getDevice());
  int cols = grad->shape().back();
  int rows = grad->shape().elements() / cols;
  auto l = rowLaunch(rows, cols);
  ROW_DISPATCH(gSoftmaxGrad, l, grad->data(), adj->data(), val->data(), rows, cols);
}

void LogSoftmaxGrad(Tensor grad, Tensor adj, Tensor val) {
  device::setDevice(adj->getDevice());
  int cols = grad->shape().back();
  int rows = grad->shape().elements() / cols;
  auto l = rowLaunch(rows, cols);
  ROW_DISPATCH(gLogSoftmaxGrad, l, grad->data(), adj->data(), val->data(), rows, cols);
}

// =============================================================================
// Cross entropy           reference: tensor_operators.cu:1115-1283
// =============================================================================
namespace {

// per-thread online statistics over a row, 128-bit loads when VEC
template <bool VEC>
__device__ __forceinline__ void rowStats(const float* __restrict__ sp, int cols, int first, int stride, float& m, float& s) {
  m = -3.4e38f;
  s = 0.f;
  if(VEC) {
    // one rescale per 128-bit chunk, ex2.approx exponentials (the reference's release build is
    // compiled with --use_fast_math): the full-precision expf made this HBM stream issue-bound.
    // Four independent 128-bit loads are in flight per thread.
    const float4* p4 = reinterpret_cast<const float4*>(sp);
    const int n4 = cols >> 2;
    auto fold = [&](const float4& q) {
      float mx = fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w));
      if(mx > m) {
        s *= __expf(m - mx);
        m = mx;
      }
      s += (__expf(q.x - m) + __expf(q.y - m)) + (__expf(q.z - m) + __expf(q.w - m));
    };
    int i = first;
    for(; i + 3 * stride < n4; i += 4 * stride) {
      float4 q0 = p4[i], q1 = p4[i + stride], q2 = p4[i + 2 * stride], q3 = p4[i + 3 * stride];
      fold(q0);
      fold(q1);
      fold(q2);
      fold(q3);
    }
    for(; i < n4; i += stride)
      fold(p4[i]);
  } else {
    for(int id = first; id < cols; id += stride)
      onlineUpdate(m, s, sp[id]);
  }
}

template <bool WARP, bool VEC>
__global__ void gCrossEntropyPick(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ pick, int rows, int cols, float* __restrict__ stats) {
  pdlEnter();
  __shared__ float smem[32];
  typedef RowCtx<WARP> R;
  for(int j = R::firstRow(); j < rows; j += R::rowStride()) {
    const float* sp = in + (size_t)j * cols;
    float m, s;
    rowStats<VEC>(sp, cols, R::firstCol(), R::colStride(), m, s);
    float M = R::max(m, smem);
    s = R::sum(s * expf(m - M), smem);
    if(R::leader()) {
      int id = (int)pick[j];
      out[j] = logf(s) - sp[id] + M;
      if(stats) {  // row max and sum of exponentials, reused by the backward pass
        stats[2 * j] = M;
        stats[2 * j + 1] = s;
      }
    }
  }
}

template <bool WARP, bool VEC>
__global__ void gCrossEntropyPickBackward(float* __restrict__ out, const float* __restrict__ adj, const float* __restrict__ in, const float* __restrict__ pick, int rows, int cols, int assign, const float* __restrict__ stats, __nv_bfloat16* __restrict__ outShadow) {
  pdlEnter();
  __shared__ float smem[32];
  typedef RowCtx<WARP> R;
  for(int j = R::firstRow(); j < rows; j += R::rowStride()) {
    const float* sp = in + (size_t)j * cols;
    float* so = out + (size_t)j * cols;
    float m, s, M;
    if(stats) {  // forward pass left the row statistics: the logits are read once, not twice
      M = stats[2 * j];
      s = stats[2 * j + 1];
    } else {
      rowStats<VEC>(sp, cols, R::firstCol(), R::colStride(), m, s);
      M = R::max(m, smem);
      s = R::sum(s * expf(m - M), smem);
    }
    int p = (int)pick[j];
    float a = adj[j];
    const float invS = 1.0f / s;  // one division per row instead of one per element
    if(VEC) {
      const float4* p4 = reinterpret_cast<const float4*>(sp);
      float4* o4 = reinterpret_cast<float4*>(so);
      int n4 = cols >> 2;
      for(int i = R::firstCol(); i < n4; i += R::colStride()) {
        float4 x = p4[i];
        float4 g = assign ? make_float4(0.f, 0.f, 0.f, 0.f) : o4[i];
        int id = i << 2;
        g.x += a * (__expf(x.x - M) * invS - (float)(id == p));
        g.y += a * (__expf(x.y - M) * invS - (float)(id + 1 == p));
        g.z += a * (__expf(x.z - M) * invS - (float)(id + 2 == p));
        g.w += a * (__expf(x.w - M) * invS - (float)(id + 3 == p));
        if(out)  // (null: shadow-only adjoint - its readers are the two products that take the bf16 copy)
          o4[i] = g;
        shadow::store4(outShadow, (size_t)j * cols + id, g);  // bf16 copy of the logits adjoint for the two products that read it
      }
    } else {
      for(int id = R::firstCol(); id < cols; id += R::colStride()) {
        float sub = (float)(id == p);
        float gval = a * (expf(sp[id] - M) * invS - sub);
        so[id] = assign ? gval : so[id] + gval;
      }
    }
  }
}

inline bool rowsVectorizable(const void* a, const void* b, int cols) {
  return cols % 4 == 0 && ((uintptr_t)a & 15) == 0 && (!b || ((uintptr_t)b & 15) == 0);
}

}  // namespace

void CrossEntropyPick(Tensor out, Tensor in, Tensor pick, Tensor stats) {
  device::setDevice(out->getDevice());
  float* statsPtr = stats ? stats->data() : nullptr;
  int cols = in->shape().back();
  int rows = in->shape().elements() / cols;
  auto l = rowLaunch(rows, cols);
  bool vec = rowsVectorizable(in->data(), nullptr, cols);
  auto st = cudaStreamOfEngine();
#define CE_FWD(W, V) launchPdl(gCrossEntropyPick<W, V>, dim3(l.grid), dim3(l.block), 0, st, out->data(), (const float*)in->data(), (const float*)pick->data(), rows, cols, statsPtr)
  if(l.warp) {
    if(vec) CE_FWD(true, true); else CE_FWD(true, false);
  } else {
    if(vec) CE_FWD(false, true); else CE_FWD(false, false);
  }
#undef CE_FWD
  CUDA_LAUNCH_CHECK();
}

void CrossEntropyPickBackward(Tensor out, Tensor adj, Tensor a, Tensor pick, Tensor stats) {
  device::setDevice(out->getDevice());
  const float* statsPtr = stats ? stats->data() : nullptr;
  int cols = out->shape().back();
  int rows = out->shape().elements() / cols;
  auto l = rowLaunch(rows, cols);
  int assign = out->takeLazyZero() ? 1 : 0;  // first writer of the logits adjoint: no memset, no read-back
  bool vec = rowsVectorizable(a->data(), out->data(), cols);
  auto st = cudaStreamOfEngine();
  __nv_bfloat16* osh = (assign && vec) ? shadow::produce(out) : nullptr;
  float* outp = shadow::fp32Target(out, osh);
#define CE_BWD(W, V) launchPdl(gCrossEntropyPickBackward<W, V>, dim3(l.grid), dim3(l.block), 0, st, outp, (const float*)adj->data(), (const float*)a->data(), (const float*)pick->data(), rows, cols, assign, statsPtr, osh)
  if(l.warp) {
    if(vec) CE_BWD(true, true); else CE_BWD(true, false);
  } else {
    if(vec) CE_BWD(false, true); else CE_BWD(false, false);
  }
#undef CE_BWD
  CUDA_LAUNCH_CHECK();
}

// =============================================================================
// Layer normalisation     reference: tensor_operators.cu:1447-1674
// =============================================================================
namespace {

template <bool WARP>
__global__ void gLNormalization(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ alpha, const float* __restrict__ beta, int rows, int cols, float eps) {
  __shared__ float smem[32];
  typedef RowCtx<WARP> R;
  for(int j = R::firstRow(); j < rows; j += R::rowStride()) {
    float* so = out + (size_t)j * cols;
    const float* sp = in + (size_t)j * cols;
    float s = 0.f;
    for(int id = R::firstCol(); id < cols; id += R::colStride())
      s += sp[id];
    float mean = R::sum(s, smem) / cols;
    float sq = 0.f;
    for(int id = R::firstCol(); id < cols; id += R::colStride()) {
      float ex = sp[id] - mean;
      sq += ex * ex;
    }
    float sigma = sqrtf(eps + (R::sum(sq, smem) / cols));  // eps inside the root, biased variance
    for(int id = R::firstCol(); id < cols; id += R::colStride()) {
      float t = alpha[id] * ((sp[id] - mean) / sigma);
      if(beta)
        t += beta[id];
      so[id] = t;
    }
  }
}

// Block per row (grid-stride over rows); thread t owns columns t, t+blockDim, ...
// (at most MAXC of them) and keeps their gamma/beta gradient sums in registers.
template <int MAXC>
__global__ void __launch_bounds__(128) gLayerNormalizationGrad(float* __restrict__ gradX,
                                                               float* __restrict__ gradGamma,
                                                               float* __restrict__ gradBeta,
                                                               const float* __restrict__ adj,
                                                               const float* __restrict__ y,
                                                               const float* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               int rows,
                                                               int cols,
                                                               float eps) {
  __shared__ float smem[32];
  float accGamma[MAXC], accBeta[MAXC];
#pragma unroll
  for(int i = 0; i < MAXC; ++i)
    accGamma[i] = accBeta[i] = 0.f;

  for(int j = blockIdx.x; j < rows; j += gridDim.x) {
    const float* xRow = x + (size_t)j * cols;
    const float* yRow = y + (size_t)j * cols;
    const float* adjRow = adj + (size_t)j * cols;
    float* gradXRow = gradX + (size_t)j * cols;

    float sum_x = 0.f, sum_adj = 0.f, sum_adj_x = 0.f;
    for(int id = threadIdx.x; id < cols; id += blockDim.x) {
      sum_x += xRow[id];
      sum_adj_x += adjRow[id] * (yRow[id] - (beta ? beta[id] : 0.f)) / gamma[id];
      sum_adj += adjRow[id];
    }
    sum_x = blockSum(sum_x, smem);
    sum_adj = blockSum(sum_adj, smem);
    sum_adj_x = blockSum(sum_adj_x, smem);
    float mean = sum_x / cols;
    float sq = 0.f;
    for(int id = threadIdx.x; id < cols; id += blockDim.x) {
      float ex = xRow[id] - mean;
      sq += ex * ex;
    }
    float sigma = sqrtf(eps + (blockSum(sq, smem) / cols));

#pragma unroll
    for(int i = 0; i < MAXC; ++i) {
      int id = threadIdx.x + i * blockDim.x;
      if(id < cols) {
        float a = adjRow[id];
        float x_hat = (yRow[id] - (beta ? beta[id] : 0.f)) / gamma[id];
        float grad_x = 0.0f;
        grad_x += cols * a;
        grad_x -= sum_adj;
        grad_x -= sum_adj_x * x_hat;
        grad_x /= (cols * sigma);
        float valX = gamma[id] * grad_x;
        float sign = (0.f < valX) - (valX < 0.f);
        valX = fabsf(valX) > 1000.f ? sign * 1000.f : valX;  // clip kept from the reference
        gradXRow[id] += valX;
        accGamma[i] += a * x_hat;
        accBeta[i] += a;
      }
    }
  }
#pragma unroll
  for(int i = 0; i < MAXC; ++i) {
    int id = threadIdx.x + i * blockDim.x;
    if(id < cols) {
      atomicAdd(gradGamma + id, accGamma[i]);
      if(beta)
        atomicAdd(gradBeta + id, accBeta[i]);
    }
  }
}

// ---- register-resident variants for rows of up to 128*VPL floats (16-byte aligned) --------
// One warp per row; a lane keeps VPL float4 of the row in registers, so every input is read
// from HBM exactly once and all row statistics are shuffle reductions (no __syncthreads).
template <int VPL>
__global__ void __launch_bounds__(256) gLNormalizationWarp(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ res, const float* __restrict__ alpha, const float* __restrict__ beta, int rows, int cols, float eps, __nv_bfloat16* __restrict__ outShadow) {
  pdlEnter();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4 g4[VPL], b4[VPL];
#pragma unroll
  for(int i = 0; i < VPL; ++i) {
    int c = (i * 32 + lane) * 4;
    g4[i] = c < cols ? *reinterpret_cast<const float4*>(alpha + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    b4[i] = (beta && c < cols) ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for(int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float* sp = in + (size_t)row * cols;
    float4 xv[VPL];
    float s = 0.f;
#pragma unroll
    for(int i = 0; i < VPL; ++i) {
      int c = (i * 32 + lane) * 4;
      xv[i] = c < cols ? *reinterpret_cast<const float4*>(sp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if(res && c < cols) {  // normalise x + residual (the "a" + "n" steps of a sub-layer in one pass)
        float4 rv = *reinterpret_cast<const float4*>(res + (size_t)row * cols + c);
        xv[i].x += rv.x;
        xv[i].y += rv.y;
        xv[i].z += rv.z;
        xv[i].w += rv.w;
      }
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
    float mean = warpSum(s) / cols;
    float sq = 0.f;
#pragma unroll
    for(int i = 0; i < VPL; ++i) {
      int c = (i * 32 + lane) * 4;
      if(c < cols) {
        float e0 = xv[i].x - mean, e1 = xv[i].y - mean, e2 = xv[i].z - mean, e3 = xv[i].w - mean;
        sq += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
      }
    }
    float sigma = sqrtf(eps + (warpSum(sq) / cols));
    float* so = out + (size_t)row * cols;
#pragma unroll
    for(int i = 0; i < VPL; ++i) {
      int c = (i * 32 + lane) * 4;
      if(c < cols) {
        float4 t;
        t.x = g4[i].x * ((xv[i].x - mean) / sigma) + b4[i].x;
        t.y = g4[i].y * ((xv[i].y - mean) / sigma) + b4[i].y;
        t.z = g4[i].z * ((xv[i].z - mean) / sigma) + b4[i].z;
        t.w = g4[i].w * ((xv[i].w - mean) / sigma) + b4[i].w;
        *reinterpret_cast<float4*>(so + c) = t;
        shadow::store4(outShadow, (size_t)row * cols + c, t);  // bf16 copy for the products that consume the output
      }
    }
  }
}

__device__ __forceinline__ float lnGradElem(float a, float x_hat, float g, float sum_adj, float sum_adj_x, float cols, float sigma) {
  float grad_x = cols * a;
  grad_x -= sum_adj;
  grad_x -= sum_adj_x * x_hat;
  grad_x /= (cols * sigma);
  float valX = g * grad_x;
  float sign = (0.f < valX) - (valX < 0.f);
  return fabsf(valX) > 1000.f ? sign * 1000.f : valX;  // clip kept from the reference
}

// WARPS = 16 (rows of up to 512 floats), one block per SM: twice the rows in flight of an 8-warp block
// at the same number of block-level column reductions (same-address red.add serialise in L2); beta is
// then re-read (L1) instead of cached, to stay within 128 registers
template <int VPL, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) gLayerNormalizationGradWarp(float* __restrict__ gradX,
                                                                   float* __restrict__ gradGamma,
                                                                   float* __restrict__ gradBeta,
                                                                   const float* __restrict__ adj,
                                                                   const float* __restrict__ y,
                                                                   const float* __restrict__ x,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   int rows,
                                                                   int cols,
                                                                   float eps,
                                                                   int assignX,
                                                                   const float* __restrict__ res,
                                                                   float* __restrict__ gradRes,
                                                                   int assignRes,
                                                                   __nv_bfloat16* __restrict__ gradXShadow) {
  pdlEnter();
  __shared__ float4 red[WARPS][32 * VPL];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4 g4[VPL], b4[VPL], accG[VPL], accB[VPL];
#pragma unroll
  for(int i = 0; i < VPL; ++i) {
    int c = (i * 32 + lane) * 4;
    g4[i] = c < cols ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    b4[i] = (beta && c < cols) ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    accG[i] = accB[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float fcols = (float)cols;
  for(int row = blockIdx.x * WARPS + warp; row < rows; row += gridDim.x * WARPS) {
    const size_t off = (size_t)row * cols;
    float4 xh[VPL], av[VPL];
    float sum_x = 0.f, sum_adj = 0.f, sum_adj_x = 0.f, sq = 0.f;
    {
      // all loads of the row are issued before the first use: ONE memory round trip per row
      float4 xv[VPL], yv[VPL], rv[VPL];
#pragma unroll
      for(int i = 0; i < VPL; ++i) {
        int c = (i * 32 + lane) * 4;
        const size_t o = off + (c < cols ? c : 0);  // clamped: the load is unconditional, the value masked below
        xv[i] = *reinterpret_cast<const float4*>(x + o);
        yv[i] = *reinterpret_cast<const float4*>(y + o);
        av[i] = *reinterpret_cast<const float4*>(adj + o);
        rv[i] = res ? *reinterpret_cast<const float4*>(res + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for(int i = 0; i < VPL; ++i) {
        int c = (i * 32 + lane) * 4;
        if(c < cols) {
          xv[i].x += rv[i].x;
          xv[i].y += rv[i].y;
          xv[i].z += rv[i].z;
          xv[i].w += rv[i].w;
          const float4 bv = WARPS > 8 ? (beta ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f)) : b4[i];
          const float4 gv = WARPS > 8 ? *reinterpret_cast<const float4*>(gamma + c) : g4[i];
          xh[i].x = (yv[i].x - bv.x) / gv.x;
          xh[i].y = (yv[i].y - bv.y) / gv.y;
          xh[i].z = (yv[i].z - bv.z) / gv.z;
          xh[i].w = (yv[i].w - bv.w) / gv.w;
        } else {
          xv[i] = xh[i] = av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        sum_x += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
        sum_adj += (av[i].x + av[i].y) + (av[i].z + av[i].w);
        sum_adj_x += (av[i].x * xh[i].x + av[i].y * xh[i].y) + (av[i].z * xh[i].z + av[i].w * xh[i].w);
      }
      sum_x = warpSum(sum_x);
      sum_adj = warpSum(sum_adj);
      sum_adj_x = warpSum(sum_adj_x);
      float mean = sum_x / fcols;
#pragma unroll
      for(int i = 0; i < VPL; ++i) {
        int c = (i * 32 + lane) * 4;
        if(c < cols) {
          float e0 = xv[i].x - mean, e1 = xv[i].y - mean, e2 = xv[i].z - mean, e3 = xv[i].w - mean;
          sq += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
        }
      }
    }
    float sigma = sqrtf(eps + (warpSum(sq) / fcols));
#pragma unroll
    for(int i = 0; i < VPL; ++i) {
      int c = (i * 32 + lane) * 4;
      if(c < cols) {
        float4 v;
        const float4 gv = WARPS > 8 ? *reinterpret_cast<const float4*>(gamma + c) : g4[i];
        v.x = lnGradElem(av[i].x, xh[i].x, gv.x, sum_adj, sum_adj_x, fcols, sigma);
        v.y = lnGradElem(av[i].y, xh[i].y, gv.y, sum_adj, sum_adj_x, fcols, sigma);
        v.z = lnGradElem(av[i].z, xh[i].z, gv.z, sum_adj, sum_adj_x, fcols, sigma);
        v.w = lnGradElem(av[i].w, xh[i].w, gv.w, sum_adj, sum_adj_x, fcols, sigma);
        if(gradRes) {  // d(x + r)/dr = 1: the residual branch receives the same gradient
          float4* gr = reinterpret_cast<float4*>(gradRes + off + c);
          float4 w = v;
          if(!assignRes) {
            float4 o = *gr;
            w.x += o.x;
            w.y += o.y;
            w.z += o.z;
            w.w += o.w;
          }
          *gr = w;
        }
        float4* gx = reinterpret_cast<float4*>(gradX + off + c);
        if(!assignX) {
          float4 o = *gx;
          v.x += o.x;
          v.y += o.y;
          v.z += o.z;
          v.w += o.w;
        }
        if(gradX)  // (null with assignX: shadow-only adjoint)
          *gx = v;
        shadow::store4(gradXShadow, off + c, v);  // only handed in when this kernel is the adjoint's one writer
        accG[i].x += av[i].x * xh[i].x;
        accG[i].y += av[i].y * xh[i].y;
        accG[i].z += av[i].z * xh[i].z;
        accG[i].w += av[i].w * xh[i].w;
        accB[i].x += av[i].x;
        accB[i].y += av[i].y;
        accB[i].z += av[i].z;
        accB[i].w += av[i].w;
      }
    }
  }
  // column sums of the block: the warps meet in shared memory, ONE atomic per block and column
#pragma unroll 1
  for(int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? gradGamma : gradBeta;
    if(!dst)
      continue;
    __syncthreads();
#pragma unroll
    for(int i = 0; i < VPL; ++i)
      red[warp][i * 32 + lane] = pass == 0 ? accG[i] : accB[i];
    __syncthreads();
    for(int e = threadIdx.x; e < 32 * VPL; e += blockDim.x) {
      int c = e * 4;
      if(c < cols) {
        float4 sum = red[0][e];
#pragma unroll
        for(int w = 1; w < WARPS; ++w) {
          float4 t = red[w][e];
          sum.x += t.x;
          sum.y += t.y;
          sum.z += t.z;
          sum.w += t.w;
        }
        redAdd4(dst + c, sum);  // gamma / beta gradients are 16-byte aligned rows (checked by the launcher)
      }
    }
  }
}

}  // namespace

bool LayerNormResidualFusable(int cols) {
  return cols % 4 == 0 && cols <= 1024;
}

void LayerNormalization(Tensor out, Tensor in, Tensor gamma, Tensor beta, float eps, Tensor residual) {
  device::setDevice(out->getDevice());
  int cols = in->shape().back();
  int rows = in->shape().elements() / cols;
  out->takeLazyZero();
  const float* bp = beta ? beta->data() : nullptr;
  const float* rp = residual ? residual->data() : nullptr;
  bool aligned = (cols % 4 == 0) && (((uintptr_t)out->data() | (uintptr_t)in->data() | (uintptr_t)gamma->data() | (uintptr_t)bp | (uintptr_t)rp) & 15) == 0;
  if(aligned && cols <= 1024) {
    int grid = std::max(1, std::min((rows + 7) / 8, kNumSMs * 4));
    auto st = cudaStreamOfEngine();
    __nv_bfloat16* osh = shadow::produce(out);
    if(cols <= 512)
      launchPdl(gLNormalizationWarp<4>, dim3(grid), dim3(256), 0, st, out->data(), (const float*)in->data(), rp, (const float*)gamma->data(), bp, rows, cols, eps, osh);
    else
      launchPdl(gLNormalizationWarp<8>, dim3(grid), dim3(256), 0, st, out->data(), (const float*)in->data(), rp, (const float*)gamma->data(), bp, rows, cols, eps, osh);
    CUDA_LAUNCH_CHECK();
    return;
  }
  ABORT_IF(residual, "LayerNormalization with a fused residual needs 16-byte aligned rows of at most 1024 floats");
  auto l = rowLaunch(rows, cols);
  ROW_DISPATCH(gLNormalization, l, out->data(), in->data(), gamma->data(), bp, rows, cols, eps);
}

void LayerNormalizationGrad(Tensor gradX, Tensor gradGamma, Tensor gradBeta, Tensor adj, Tensor y, Tensor x, Tensor gamma, Tensor beta, float eps, Tensor residual, Tensor gradResidual) {
  device::setDevice(adj->getDevice());
  int cols = y->shape().back();
  int rows = y->shape().elements() / cols;
  auto st = cudaStreamOfEngine();
  {
    const float* bp = beta ? beta->data() : nullptr;
    float* gbp = gradBeta ? gradBeta->data() : nullptr;
    bool aligned = (cols % 4 == 0)
                   && (((uintptr_t)gradX->memory()->data() | (uintptr_t)adj->data() | (uintptr_t)y->data() | (uintptr_t)x->data() | (uintptr_t)gamma->data() | (uintptr_t)bp) & 15) == 0;
    const float* rp = residual ? residual->data() : nullptr;
    aligned = aligned && (((uintptr_t)rp | (uintptr_t)(gradResidual ? gradResidual->memory()->data() : nullptr) | (uintptr_t)gradGamma->data() | (uintptr_t)gbp) & 15) == 0;
    if(aligned && cols <= 1024) {
      int assignX = gradX->takeLazyZero() ? 1 : 0;
      int assignRes = (gradResidual && gradResidual->takeLazyZero()) ? 1 : 0;
      float* grp = gradResidual ? gradResidual->data() : nullptr;
      __nv_bfloat16* gxs = assignX ? shadow::produce(gradX) : nullptr;
      float* gxp = assignX ? shadow::fp32Target(gradX, gxs) : gradX->data();
      // few, fat blocks: every block ends with one 128-bit reduction per 4 columns for gamma and
      // beta; same-address reductions serialise in L2, so one block per SM is the sweet spot
      if(cols <= 512) {
        int grid = std::max(1, std::min((rows + 15) / 16, kNumSMs));
        launchPdl(gLayerNormalizationGradWarp<4, 16>, dim3(grid), dim3(512), 0, st, gxp, gradGamma->data(), gbp, (const float*)adj->data(), (const float*)y->data(), (const float*)x->data(),
                  (const float*)gamma->data(), bp, rows, cols, eps, assignX, rp, grp, assignRes, gxs);
      } else {
        int grid = std::max(1, std::min((rows + 15) / 16, kNumSMs));
        launchPdl(gLayerNormalizationGradWarp<8, 8>, dim3(grid), dim3(256), 0, st, gxp, gradGamma->data(), gbp, (const float*)adj->data(), (const float*)y->data(), (const float*)x->data(),
                  (const float*)gamma->data(), bp, rows, cols, eps, assignX, rp, grp, assignRes, gxs);
      }
      CUDA_LAUNCH_CHECK();
      return;
    }
  }
  ABORT_IF(residual || gradResidual, "LayerNormalizationGrad with a fused residual needs 16-byte aligned rows of at most 1024 floats");
  int grid = std::max(1, std::min(rows, kNumSMs * 4));
#define LN_BWD(M)                                                                                                                  \
  gLayerNormalizationGrad<M><<<grid, 128, 0, st>>>(gradX->data(), gradGamma->data(), gradBeta ? gradBeta->data() : nullptr, adj->data(), \
                                                   y->data(), x->data(), gamma->data(), beta ? beta->data() : nullptr, rows, cols, eps)
  if(cols <= 128 * 4)
    LN_BWD(4);
  else if(cols <= 128 * 8)
    LN_BWD(8);
  else if(cols <= 128 * 32)
    LN_BWD(32);
  else
    ABORT("LayerNormalizationGrad: rows longer than 4096 are not supported", cols);
#undef LN_BWD
  CUDA_LAUNCH_CHECK();
}

// =============================================================================
// Bahdanau attention score   reference: tensor_operators.cu:1307-1445
// =============================================================================
namespace {

template <bool WARP>
__global__ void gAtt(float* __restrict__ out, const float* __restrict__ va, const float* __restrict__ ctx, const float* __restrict__ state, int m, int k, int b, int t) {
  __shared__ float smem[32];
  typedef RowCtx<WARP> R;
  for(int j = R::firstRow(); j < m; j += R::rowStride()) {
    const float* ctxRow = ctx + (size_t)(j % (b * t)) * k;
    const float* stateRow = state + (size_t)((j / (b * t)) * b + j % b) * k;
    float s = 0.f;
    for(int id = R::firstCol(); id < k; id += R::colStride()) {
      float z = ctxRow[id] + stateRow[id];
      s += tanhf(z) * va[id];
    }
    s = R::sum(s, smem);
    if(R::leader())
      out[j] = s;
  }
}

// 16-byte aligned rows, k % 4 == 0: one warp per (source position, sentence) row, 128-bit loads, four
// independent partial sums per lane (the block-per-row form above spends two block barriers per row
// of 2048 elements: 22 us per decoder step at 50 x 64 rows).
__global__ void __launch_bounds__(256) gAttVec(float* __restrict__ out, const float* __restrict__ va, const float* __restrict__ ctx, const float* __restrict__ state, int m, int k, int b, int t) {
  const int lane = threadIdx.x & 31;
  const int k4 = k >> 2;
  const float4* va4 = reinterpret_cast<const float4*>(va);
  for(int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < m; j += gridDim.x * (blockDim.x >> 5)) {
    const float4* ctxRow = reinterpret_cast<const float4*>(ctx + (size_t)(j % (b * t)) * k);
    const float4* stateRow = reinterpret_cast<const float4*>(state + (size_t)((j / (b * t)) * b + j % b) * k);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
    for(int id = lane; id < k4; id += 32) {
      const float4 c = ctxRow[id], st = stateRow[id], v = va4[id];
      s0 += tanhf(c.x + st.x) * v.x;
      s1 += tanhf(c.y + st.y) * v.y;
      s2 += tanhf(c.z + st.z) * v.z;
      s3 += tanhf(c.w + st.w) * v.w;
    }
    float s = warpSum((s0 + s1) + (s2 + s3));
    if(lane == 0)
      out[j] = s;
  }
}

// grid = (batch n, column chunks); a thread owns ONE column c of ONE batch
// element and walks the rows j = bIdx, bIdx+n, ... (time steps): state row and
// va are loaded once, gState / gVa sums stay in registers.
__global__ void __launch_bounds__(256) gAttBack(float* __restrict__ gVa,
                                                float* __restrict__ gContext,
                                                float* __restrict__ gState,
                                                const float* __restrict__ va,
                                                const float* __restrict__ context,
                                                const float* __restrict__ state,
                                                const float* __restrict__ adj,
                                                int m,
                                                int k,
                                                int n,
                                                int assignState,
                                                __nv_bfloat16* __restrict__ gStateShadow) {
  int c = blockIdx.y * blockDim.x + threadIdx.x;
  int bIdx = blockIdx.x;
  if(c >= k)
    return;
  float s = state[(size_t)bIdx * k + c];
  float v = va[c];
  float accState = 0.f, accVa = 0.f;
  // five time steps per round: all fifteen loads are issued before the first use (one memory round trip per
  // round instead of one per step - the walk is a serial chain of 50 otherwise); same summation order as before
  constexpr int U = 5;
  int j = bIdx;
  for(; j + (U - 1) * n < m; j += U * n) {
    float cx[U], a[U], g[U];
#pragma unroll
    for(int u = 0; u < U; ++u) {
      const size_t at = (size_t)(j + u * n) * k + c;
      cx[u] = context[at];
      g[u] = gContext[at];
      a[u] = adj[j + u * n];
    }
#pragma unroll
    for(int u = 0; u < U; ++u) {
      float t = tanhf(cx[u] + s);
      float r = v * (1.f - t * t);
      gContext[(size_t)(j + u * n) * k + c] = g[u] + r * a[u];
      accState += r * a[u];
      accVa += t * a[u];
    }
  }
  for(; j < m; j += n) {
    float z = context[(size_t)j * k + c] + s;
    float t = tanhf(z);
    float r = v * (1.f - t * t);
    float a = adj[j];
    gContext[(size_t)j * k + c] += r * a;
    accState += r * a;
    accVa += t * a;
  }
  // a lazily-zero state adjoint (one attention per decoder step writes it) is assigned, and leaves the bf16 copy its two
  // backward products read
  const float gs = assignState ? accState : gState[(size_t)bIdx * k + c] + accState;
  gState[(size_t)bIdx * k + c] = gs;
  shadow::store1(gStateShadow, (size_t)bIdx * k + c, gs);
  atomicAdd(gVa + c, accVa);
}

}  // namespace

void Att(Tensor out, Tensor va, Tensor context, Tensor state) {
  device::setDevice(out->getDevice());
  int m = out->shape().elements() / out->shape().back();
  int k = context->shape()[-1];
  int b = context->shape()[-2];
  int t = context->shape()[-3];
  if(k % 4 == 0 && k >= 128 && ((((uintptr_t)va->data()) | ((uintptr_t)context->data()) | ((uintptr_t)state->data())) & 15) == 0) {
    int blocks = std::max(1, std::min((m + 7) / 8, kNumSMs * 8));
    gAttVec<<<blocks, 256, 0, cudaStreamOfEngine()>>>(out->data(), va->data(), context->data(), state->data(), m, k, b, t);
    CUDA_LAUNCH_CHECK();
    return;
  }
  auto l = rowLaunch(m, k);
  ROW_DISPATCH(gAtt, l, out->data(), va->data(), context->data(), state->data(), m, k, b, t);
}

void AttBack(Tensor gVa, Tensor gContext, Tensor gState, Tensor va, Tensor context, Tensor state, Tensor adj) {
  device::setDevice(adj->getDevice());
  int m = adj->shape().elements() / adj->shape().back();
  int k = context->shape()[-1];
  int n = context->shape()[-2];
  dim3 grid(n, (k + 255) / 256);
  // (the kernel covers every element of gState exactly when the state has one row per sentence: no beam dimension)
  const bool whole = (size_t)gState->size() == (size_t)n * k;
  const int assignState = (whole && gState->takeLazyZero()) ? 1 : 0;
  __nv_bfloat16* gss = assignState ? shadow::produce(gState) : nullptr;
  gAttBack<<<grid, 256, 0, cudaStreamOfEngine()>>>(
      gVa->data(), gContext->data(), gState->data(), va->data(), context->data(), state->data(), adj->data(), m, k, n, assignState, gss);
  CUDA_LAUNCH_CHECK();
}

// =============================================================================
// GRU / LSTM fused cells      reference: tensor_operators.cu:934-1113, 1749-2031
// =============================================================================
namespace {

__global__ void __launch_bounds__(256) gGRUFastForward(float* __restrict__ out,
                                                       const float* __restrict__ state,
                                                       const float* __restrict__ xW,
                                                       const float* __restrict__ sU,
                                                       const float* __restrict__ b,
                                                       const float* __restrict__ mask,
                                                       int rows,
                                                       int cols,
                                                       bool final,
                                                       __nv_bfloat16* __restrict__ outShadow) {
  long long n = (long long)rows * cols;
  for(long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    int j = (int)(idx / cols);
    int i = (int)(idx - (long long)j * cols);
    float m = !mask || mask[j];
    const float* xWrow = xW + (size_t)j * cols * 3;
    const float* sUrow = sU + (size_t)j * cols * 3;
    float st = state[idx];

    float r = stableLogit(xWrow[i] + sUrow[i] + b[i]);
    int k = i + cols;
    float z = stableLogit(xWrow[k] + sUrow[k] + b[k]);
    int l = i + 2 * cols;
    float h;
    if(final)
      h = tanhf(xWrow[l] + (sUrow[l] + b[l]) * r);
    else
      h = tanhf(xWrow[l] + sUrow[l] * r + b[l]);
    float o = (1.0f - z) * h + z * st;
    float res = m * o + (1 - m) * st;
    out[idx] = res;
    shadow::store1(outShadow, (size_t)idx, res);  // the next step's state product reads the bf16 copy (BF16S mode)
  }
}

// grid = (column blocks, row splits): a thread owns one column i and walks the
// rows of its split; the three bias-gradient sums stay in registers.
__global__ void __launch_bounds__(128) gGRUFastBackward(float* __restrict__ outState,
                                                        float* __restrict__ outXW,
                                                        float* __restrict__ outSU,
                                                        float* __restrict__ outB,
                                                        const float* __restrict__ state,
                                                        const float* __restrict__ xW,
                                                        const float* __restrict__ sU,
                                                        const float* __restrict__ b,
                                                        const float* __restrict__ mask,
                                                        const float* __restrict__ adj,
                                                        int rows,
                                                        int cols,
                                                        int rowsPerSplit,
                                                        bool final) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= cols)
    return;
  int k = i + cols;
  int l = i + 2 * cols;
  float bi = b[i], bk = b[k], bl = b[l];
  float accR = 0.f, accZ = 0.f, accX = 0.f;
  int j0 = blockIdx.y * rowsPerSplit;
  int j1 = min(rows, j0 + rowsPerSplit);
  for(int j = j0; j < j1; ++j) {
    float m = !mask || mask[j];
    const float* rowXW = xW + (size_t)j * cols * 3;
    const float* rowSU = sU + (size_t)j * cols * 3;
    float st = state[(size_t)j * cols + i];

    float r = stableLogit(rowXW[i] + rowSU[i] + bi);
    float z = stableLogit(rowXW[k] + rowSU[k] + bk);
    float h;
    if(final)
      h = tanhf(rowXW[l] + (rowSU[l] + bl) * r);
    else
      h = tanhf(rowXW[l] + rowSU[l] * r + bl);

    float a = adj[(size_t)j * cols + i];
    float t = (1 - z) * (1 - h * h);

    if(outState)
      outState[(size_t)j * cols + i] += (m * z - m + 1) * a;

    float dfdxW_r = m * r * (1 - r) * t * a;
    if(final)
      dfdxW_r *= rowSU[l] + bl;
    else
      dfdxW_r *= rowSU[l];
    float dfdxW_z = m * (1 - z) * z * (st - h) * a;
    float dfdxW_x = m * t * a;

    if(outXW) {
      float* o = outXW + (size_t)j * cols * 3;
      o[i] += dfdxW_r;
      o[k] += dfdxW_z;
      o[l] += dfdxW_x;
    }
    if(outSU) {
      float* o = outSU + (size_t)j * cols * 3;
      o[i] += dfdxW_r;
      o[k] += dfdxW_z;
      o[l] += dfdxW_x * r;
    }
    accR += dfdxW_r;
    accZ += dfdxW_z;
    accX += final ? dfdxW_x * r : dfdxW_x;
  }
  if(outB) {
    atomicAdd(outB + i, accR);
    atomicAdd(outB + k, accZ);
    atomicAdd(outB + l, accX);
  }
}

// Vectorised backward of the fused GRU gate (cols % 4 == 0, 16-byte aligned rows): a lane owns FOUR
// adjacent columns, a warp one row at a time, the 8 warps of a block walk `rowsPerBlock` rows of the
// same 128-column strip.  Every input is read once with 128-bit loads that are all issued before
// the first use; outputs flagged in `assignMask` (lazily-zero adjoints, tensors/tensor.h) are
// stored instead of read-modify-written; the three bias-gradient sums meet in shared memory and
// leave the block as one atomic per column.  (The scalar kernel above walked the rows serially:
// 22 us for a 64 x 1024 state, the longest kernel of the recurrent backward chain.)
__global__ void __launch_bounds__(256) gGRUFastBackwardVec(float* __restrict__ outState,
                                                           float* __restrict__ outXW,
                                                           float* __restrict__ outSU,
                                                           float* __restrict__ outB,
                                                           const float* __restrict__ state,
                                                           const float* __restrict__ xW,
                                                           const float* __restrict__ sU,
                                                           const float* __restrict__ b,
                                                           const float* __restrict__ mask,
                                                           const float* __restrict__ adj,
                                                           int rows,
                                                           int cols,
                                                           int rowsPerBlock,
                                                           int assignMask,
                                                           bool final,
                                                           __nv_bfloat16* __restrict__ shadowSU,
                                                           __nv_bfloat16* __restrict__ shadowXW) {
  __shared__ float4 red[3][8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = (blockIdx.x * 32 + lane) * 4;
  const bool active = i < cols;
  float accR[4] = {0.f, 0.f, 0.f, 0.f}, accZ[4] = {0.f, 0.f, 0.f, 0.f}, accX[4] = {0.f, 0.f, 0.f, 0.f};
  if(active) {
    const float4 b4r = *(const float4*)(b + i), b4z = *(const float4*)(b + cols + i), b4x = *(const float4*)(b + 2 * cols + i);
    const float br[4] = {b4r.x, b4r.y, b4r.z, b4r.w}, bz[4] = {b4z.x, b4z.y, b4z.z, b4z.w}, bx[4] = {b4x.x, b4x.y, b4x.z, b4x.w};
    const int j0 = blockIdx.y * rowsPerBlock;
    const int j1 = min(rows, j0 + rowsPerBlock);
    for(int j = j0 + warp; j < j1; j += 8) {
      const size_t rs = (size_t)j * cols + i, rg = (size_t)j * cols * 3 + i;
      const float4 st4 = *(const float4*)(state + rs), a4 = *(const float4*)(adj + rs);
      const float4 xr4 = *(const float4*)(xW + rg), xz4 = *(const float4*)(xW + rg + cols), xx4 = *(const float4*)(xW + rg + 2 * cols);
      const float4 sr4 = *(const float4*)(sU + rg), sz4 = *(const float4*)(sU + rg + cols), sx4 = *(const float4*)(sU + rg + 2 * cols);
      float4 oS = make_float4(0.f, 0.f, 0.f, 0.f), oXr = oS, oXz = oS, oXx = oS, oUr = oS, oUz = oS, oUx = oS;
      if(outState && !(assignMask & 1))
        oS = *(const float4*)(outState + rs);
      if(outXW && !(assignMask & 2)) {
        oXr = *(const float4*)(outXW + rg);
        oXz = *(const float4*)(outXW + rg + cols);
        oXx = *(const float4*)(outXW + rg + 2 * cols);
      }
      if(outSU && !(assignMask & 4)) {
        oUr = *(const float4*)(outSU + rg);
        oUz = *(const float4*)(outSU + rg + cols);
        oUx = *(const float4*)(outSU + rg + 2 * cols);
      }
      const float m = !mask || mask[j];
      const float st[4] = {st4.x, st4.y, st4.z, st4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float xr[4] = {xr4.x, xr4.y, xr4.z, xr4.w}, xz[4] = {xz4.x, xz4.y, xz4.z, xz4.w}, xx[4] = {xx4.x, xx4.y, xx4.z, xx4.w};
      const float sr[4] = {sr4.x, sr4.y, sr4.z, sr4.w}, sz[4] = {sz4.x, sz4.y, sz4.z, sz4.w}, sx[4] = {sx4.x, sx4.y, sx4.z, sx4.w};
      float dS[4], dR[4], dZ[4], dX[4], dXu[4];
#pragma unroll
      for(int e = 0; e < 4; ++e) {
        float r = stableLogit(xr[e] + sr[e] + br[e]);
        float z = stableLogit(xz[e] + sz[e] + bz[e]);
        float h;
        if(final)
          h = tanhf(xx[e] + (sx[e] + bx[e]) * r);
        else
          h = tanhf(xx[e] + sx[e] * r + bx[e]);
        float t = (1 - z) * (1 - h * h);
        dS[e] = (m * z - m + 1) * a[e];
        float dfdxW_r = m * r * (1 - r) * t * a[e];
        if(final)
          dfdxW_r *= sx[e] + bx[e];
        else
          dfdxW_r *= sx[e];
        float dfdxW_z = m * (1 - z) * z * (st[e] - h) * a[e];
        float dfdxW_x = m * t * a[e];
        dR[e] = dfdxW_r;
        dZ[e] = dfdxW_z;
        dX[e] = dfdxW_x;
        dXu[e] = dfdxW_x * r;
        accR[e] += dfdxW_r;
        accZ[e] += dfdxW_z;
        accX[e] += final ? dfdxW_x * r : dfdxW_x;
      }
      if(outState)
        *(float4*)(outState + rs) = make_float4(oS.x + dS[0], oS.y + dS[1], oS.z + dS[2], oS.w + dS[3]);
      if(outXW) {
        *(float4*)(outXW + rg) = make_float4(oXr.x + dR[0], oXr.y + dR[1], oXr.z + dR[2], oXr.w + dR[3]);
        *(float4*)(outXW + rg + cols) = make_float4(oXz.x + dZ[0], oXz.y + dZ[1], oXz.z + dZ[2], oXz.w + dZ[3]);
        *(float4*)(outXW + rg + 2 * cols) = make_float4(oXx.x + dX[0], oXx.y + dX[1], oXx.z + dX[2], oXx.w + dX[3]);
        shadow::store4(shadowXW, rg, make_float4(dR[0], dR[1], dR[2], dR[3]));  // (only handed in for an assigned adjoint)
        shadow::store4(shadowXW, rg + cols, make_float4(dZ[0], dZ[1], dZ[2], dZ[3]));
        shadow::store4(shadowXW, rg + 2 * cols, make_float4(dX[0], dX[1], dX[2], dX[3]));
      }
      if(outSU) {
        *(float4*)(outSU + rg) = make_float4(oUr.x + dR[0], oUr.y + dR[1], oUr.z + dR[2], oUr.w + dR[3]);
        *(float4*)(outSU + rg + cols) = make_float4(oUz.x + dZ[0], oUz.y + dZ[1], oUz.z + dZ[2], oUz.w + dZ[3]);
        *(float4*)(outSU + rg + 2 * cols) = make_float4(oUx.x + dXu[0], oUx.y + dXu[1], oUx.z + dXu[2], oUx.w + dXu[3]);
        // assigned adjoint with one writer: its two backward products read this bf16 copy (BF16S mode)
        shadow::store4(shadowSU, rg, make_float4(dR[0], dR[1], dR[2], dR[3]));
        shadow::store4(shadowSU, rg + cols, make_float4(dZ[0], dZ[1], dZ[2], dZ[3]));
        shadow::store4(shadowSU, rg + 2 * cols, make_float4(dXu[0], dXu[1], dXu[2], dXu[3]));
      }
    }
  }
  if(!outB)
    return;
  red[0][warp][lane] = make_float4(accR[0], accR[1], accR[2], accR[3]);
  red[1][warp][lane] = make_float4(accZ[0], accZ[1], accZ[2], accZ[3]);
  red[2][warp][lane] = make_float4(accX[0], accX[1], accX[2], accX[3]);
  __syncthreads();
  // 3 gates x 128 columns = 384 sums of 8 partials: thread t takes gate t / 128, column t % 128 (and 256.. a second one)
  for(int t = threadIdx.x; t < 384; t += 256) {
    const int g = t >> 7, c = t & 127;
    const int col = blockIdx.x * 128 + c;
    if(col >= cols)
      continue;
    float s = 0.f;
#pragma unroll
    for(int w = 0; w < 8; ++w)
      s += ((const float*)&red[g][w][c >> 2])[c & 3];
    atomicAdd(outB + g * cols + col, s);
  }
}

__global__ void __launch_bounds__(256) gLSTMCellForward(float* __restrict__ out,
                                                        const float* __restrict__ cell,
                                                        const float* __restrict__ xW,
                                                        const float* __restrict__ sU,
                                                        const float* __restrict__ b,
                                                        const float* __restrict__ mask,
                                                        int rows,
                                                        int cols) {
  long long n = (long long)rows * cols;
  for(long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    int j = (int)(idx / cols);
    int i = (int)(idx - (long long)j * cols);
    float m = !mask || mask[j];
    const float* xWrow = xW + (size_t)j * cols * 4;
    const float* sUrow = sU + (size_t)j * cols * 4;
    float c = cell[idx];
    float gf = stableLogit(xWrow[i] + sUrow[i] + b[i]);
    int k = i + cols;
    float gi = stableLogit(xWrow[k] + sUrow[k] + b[k]);
    int l = i + 2 * cols;
    float gc = tanhf(xWrow[l] + sUrow[l] + b[l]);
    float cout = gf * c + gi * gc;
    out[idx] = m * cout + (1 - m) * c;
  }
}

__global__ void __launch_bounds__(256) gLSTMOutputForward(float* __restrict__ out,
                                                          const float* __restrict__ cell,
                                                          const float* __restrict__ xW,
                                                          const float* __restrict__ sU,
                                                          const float* __restrict__ b,
                                                          int rows,
                                                          int cols) {
  long long n = (long long)rows * cols;
  for(long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    int j = (int)(idx / cols);
    int i = (int)(idx - (long long)j * cols);
    const float* xWrow = xW + (size_t)j * cols * 4;
    const float* sUrow = sU + (size_t)j * cols * 4;
    int k = i + 3 * cols;
    float go = stableLogit(xWrow[k] + sUrow[k] + b[k]);
    out[idx] = go * tanhf(cell[idx]);
  }
}

__global__ void __launch_bounds__(128) gLSTMCellBackward(float* __restrict__ outCell,
                                                         float* __restrict__ outXW,
                                                         float* __restrict__ outSU,
                                                         float* __restrict__ outB,
                                                         const float* __restrict__ cell,
                                                         const float* __restrict__ xW,
                                                         const float* __restrict__ sU,
                                                         const float* __restrict__ b,
                                                         const float* __restrict__ mask,
                                                         const float* __restrict__ adj,
                                                         int rows,
                                                         int cols,
                                                         int rowsPerSplit) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= cols)
    return;
  int k = i + cols;
  int l = i + 2 * cols;
  float bi = b[i], bk = b[k], bl = b[l];
  float accF = 0.f, accI = 0.f, accC = 0.f;
  int j0 = blockIdx.y * rowsPerSplit;
  int j1 = min(rows, j0 + rowsPerSplit);
  for(int j = j0; j < j1; ++j) {
    float m = !mask || mask[j];
    const float* xWrow = xW + (size_t)j * cols * 4;
    const float* sUrow = sU + (size_t)j * cols * 4;
    float c = cell[(size_t)j * cols + i];
    float gf = stableLogit(xWrow[i] + sUrow[i] + bi);
    float gi = stableLogit(xWrow[k] + sUrow[k] + bk);
    float gc = tanhf(xWrow[l] + sUrow[l] + bl);
    float a = adj[(size_t)j * cols + i];

    if(outCell)
      outCell[(size_t)j * cols + i] += (m * gf - m + 1) * a;

    float dcdxf = m * c * gf * (1 - gf) * a;
    float dcdb_i = m * gc * gi * (1 - gi) * a;
    float dcdxc = m * gi * (1 - gc * gc) * a;
    if(outXW) {
      float* o = outXW + (size_t)j * cols * 4;
      o[i] += dcdxf;
      o[k] += dcdb_i;
      o[l] += dcdxc;
    }
    if(outSU) {
      float* o = outSU + (size_t)j * cols * 4;
      o[i] += dcdxf;
      o[k] += dcdb_i;
      o[l] += dcdxc;
    }
    accF += dcdxf;
    accI += dcdb_i;
    accC += dcdxc;
  }
  if(outB) {
    atomicAdd(outB + i, accF);
    atomicAdd(outB + k, accI);
    atomicAdd(outB + l, accC);
  }
}

__global__ void __launch_bounds__(128) gLSTMOutputBackward(float* __restrict__ outCell,
                                                           float* __restrict__ outXW,
                                                           float* __restrict__ outSU,
                                                           float* __restrict__ outB,
                                                           const float* __restrict__ cell,
                                                           const float* __restrict__ xW,
                                                           const float* __restrict__ sU,
                                                           const float* __restrict__ b,
                                                           const float* __restrict__ adj,
                                                           int rows,
                                                           int cols,
                                                           int rowsPerSplit) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= cols)
    return;
  int k = i + 3 * cols;
  float bk = b[k];
  float accO = 0.f;
  int j0 = blockIdx.y * rowsPerSplit;
  int j1 = min(rows, j0 + rowsPerSplit);
  for(int j = j0; j < j1; ++j) {
    const float* xWrow = xW + (size_t)j * cols * 4;
    const float* sUrow = sU + (size_t)j * cols * 4;
    float go = stableLogit(xWrow[k] + sUrow[k] + bk);
    float t = tanhf(cell[(size_t)j * cols + i]);
    float a = adj[(size_t)j * cols + i];
    if(outCell)
      outCell[(size_t)j * cols + i] += go * (1 - t * t) * a;
    float dcdxo = t * go * (1 - go) * a;
    if(outXW)
      outXW[(size_t)j * cols * 4 + k] += dcdxo;
    if(outSU)
      outSU[(size_t)j * cols * 4 + k] += dcdxo;
    accO += dcdxo;
  }
  if(outB)
    atomicAdd(outB + k, accO);
}

struct CellBwdLaunch {
  dim3 grid;
  int rowsPerSplit;
};
inline CellBwdLaunch cellBwdLaunch(int rows, int cols) {
  int colBlocks = (cols + 127) / 128;
  // aim for >= 2 blocks per SM; never fewer than 4 rows per split
  int splits = std::max(1, std::min((kNumSMs * 2 + colBlocks - 1) / colBlocks, (rows + 3) / 4));
  CellBwdLaunch l;
  l.rowsPerSplit = (rows + splits - 1) / splits;
  splits = (rows + l.rowsPerSplit - 1) / l.rowsPerSplit;
  l.grid = dim3(colBlocks, splits);
  return l;
}

inline float* dataOrNull(const std::vector<Tensor>& v, size_t i) {
  return (i < v.size() && v[i]) ? v[i]->data() : nullptr;
}

}  // namespace

void GRUFastForward(Tensor out, std::vector<Tensor> inputs, bool final) {
  device::setDevice(out->getDevice());
  int cols = out->shape().back();
  int rows = out->shape().elements() / cols;
  out->takeLazyZero();
  gGRUFastForward<<<gridFor((size_t)rows * cols, 256), 256, 0, cudaStreamOfEngine()>>>(
      out->data(), inputs[0]->data(), inputs[1]->data(), inputs[2]->data(), inputs[3]->data(), dataOrNull(inputs, 4), rows, cols, final, shadow::produce(out));
  CUDA_LAUNCH_CHECK();
}

void GRUFastBackward(std::vector<Tensor> outputs, std::vector<Tensor> inputs, Tensor adj, bool final) {
  device::setDevice(adj->getDevice());
  int cols = adj->shape().back();
  int rows = adj->shape().elements() / cols;
  static const bool scalarOnly = std::getenv("MRN_GRU_SCALAR_BACKWARD") != nullptr;
  bool vec = !scalarOnly && cols % 4 == 0;
  for(size_t k = 0; vec && k < 4; ++k)
    vec = ((uintptr_t)inputs[k]->data() & 15) == 0;
  vec = vec && ((uintptr_t)adj->data() & 15) == 0;
  for(size_t k = 0; vec && k < 3 && k < outputs.size(); ++k)
    if(outputs[k])
      vec = ((uintptr_t)outputs[k]->rawData() & 15) == 0;
  if(vec) {
    // lazily-zero adjoints (the first writer assigns): no memset, no read of the old value
    int assignMask = 0;
    float* out[3] = {nullptr, nullptr, nullptr};
    for(size_t k = 0; k < 3 && k < outputs.size(); ++k)
      if(outputs[k]) {
        if(outputs[k]->takeLazyZero())
          assignMask |= 1 << k;
        out[k] = outputs[k]->rawData();
      }
    __nv_bfloat16* shadowSU = (out[2] && (assignMask & 4)) ? shadow::produce(outputs[2]) : nullptr;
    __nv_bfloat16* shadowXW = (out[1] && (assignMask & 2)) ? shadow::produce(outputs[1]) : nullptr;  // per-step input projections (conditional cell)
    int strips = (cols / 4 + 31) / 32;
    // rows of a strip are split over blocks until the grid has about two blocks per SM, 8 rows (one per warp) at least
    int splits = std::max(1, std::min((kNumSMs * 2 + strips - 1) / strips, (rows + 7) / 8));
    int rowsPerBlock = ((rows + splits - 1) / splits + 7) / 8 * 8;
    splits = (rows + rowsPerBlock - 1) / rowsPerBlock;
    gGRUFastBackwardVec<<<dim3(strips, splits), 256, 0, cudaStreamOfEngine()>>>(out[0], out[1], out[2], dataOrNull(outputs, 3), inputs[0]->data(), inputs[1]->data(), inputs[2]->data(),
                                                                                 inputs[3]->data(), dataOrNull(inputs, 4), adj->data(), rows, cols, rowsPerBlock, assignMask, final, shadowSU, shadowXW);
    CUDA_LAUNCH_CHECK();
    return;
  }
  auto l = cellBwdLaunch(rows, cols);
  gGRUFastBackward<<<l.grid, 128, 0, cudaStreamOfEngine()>>>(dataOrNull(outputs, 0),
                                                            dataOrNull(outputs, 1),
                                                            dataOrNull(outputs, 2),
                                                            dataOrNull(outputs, 3),
                                                            inputs[0]->data(),
                                                            inputs[1]->data(),
                                                            inputs[2]->data(),
                                                            inputs[3]->data(),
                                                            dataOrNull(inputs, 4),
                                                            adj->data(),
                                                            rows,
                                                            cols,
                                                            l.rowsPerSplit,
                                                            final);
  CUDA_LAUNCH_CHECK();
}

void LSTMCellForward(Tensor out, std::vector<Tensor> inputs) {
  device::setDevice(out->getDevice());
  int cols = out->shape().back();
  int rows = out->shape().elements() / cols;
  gLSTMCellForward<<<gridFor((size_t)rows * cols, 256), 256, 0, cudaStreamOfEngine()>>>(
      out->data(), inputs[0]->data(), inputs[1]->data(), inputs[2]->data(), inputs[3]->data(), dataOrNull(inputs, 4), rows, cols);
  CUDA_LAUNCH_CHECK();
}

void LSTMOutputForward(Tensor out, std::vector<Tensor> inputs) {
  device::setDevice(out->getDevice());
  int cols = out->shape().back();
  int rows = out->shape().elements() / cols;
  gLSTMOutputForward<<<gridFor((size_t)rows * cols, 256), 256, 0, cudaStreamOfEngine()>>>(
      out->data(), inputs[0]->data(), inputs[1]->data(), inputs[2]->data(), inputs[3]->data(), rows, cols);
  CUDA_LAUNCH_CHECK();
}

void LSTMCellBackward(std::vector<Tensor> outputs, std::vector<Tensor> inputs, Tensor adj) {
  device::setDevice(adj->getDevice());
  int cols = adj->shape().back();
  int rows = adj->shape().elements() / cols;
  auto l = cellBwdLaunch(rows, cols);
  gLSTMCellBackward<<<l.grid, 128, 0, cudaStreamOfEngine()>>>(dataOrNull(outputs, 0),
                                                             dataOrNull(outputs, 1),
                                                             dataOrNull(outputs, 2),
                                                             dataOrNull(outputs, 3),
                                                             inputs[0]->data(),
                                                             inputs[1]->data(),
                                                             inputs[2]->data(),
                                                             inputs[3]->data(),
                                                             dataOrNull(inputs, 4),
                                                             adj->data(),
                                                             rows,
                                                             cols,
                                                             l.rowsPerSplit);
  CUDA_LAUNCH_CHECK();
}

void LSTMOutputBackward(std::vector<Tensor> outputs, std::vector<Tensor> inputs, Tensor adj) {
  device::setDevice(adj->getDevice());
  int cols = adj->shape().back();
  int rows = adj->shape().elements() / cols;
  auto l = cellBwdLaunch(rows, cols);
  gLSTMOutputBackward<<<l.grid, 128, 0, cudaStreamOfEngine()>>>(dataOrNull(outputs, 0),
                                                               dataOrNull(outputs, 1),
                                                               dataOrNull(outputs, 2),
                                                               dataOrNull(outputs, 3),
                                                               inputs[0]->data(),
                                                               inputs[1]->data(),
                                                               inputs[2]->data(),
                                                               inputs[3]->data(),
                                                               adj->data(),
                                                               rows,
                                                               cols,
                                                               l.rowsPerSplit);
  CUDA_LAUNCH_CHECK();
}

// =============================================================================
// Highway                    reference: tensor_operators.cu:2033-2104
// =============================================================================
namespace {
__global__ void gHighwayForward(float* out, const float* in1, const float* in2, const float* t, size_t length) {
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < length; i += (size_t)gridDim.x * blockDim.x) {
    float sigma = stableLogit(t[i]);
    out[i] = in1[i] * sigma + in2[i] * (1.f - sigma);
  }
}
__global__ void gHighwayBackward(float* out1, float* out2, float* outt, const float* in1, const float* in2, const float* t, const float* adj, size_t length) {
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < length; i += (size_t)gridDim.x * blockDim.x) {
    float sigma = stableLogit(t[i]);
    out1[i] = sigma * adj[i];  // ASSIGNS, as the reference (:2074-2077)
    out2[i] = (1.f - sigma) * adj[i];
    outt[i] = sigma * (1.f - sigma) * (in1[i] - in2[i]) * adj[i];
  }
}
}  // namespace

void HighwayForward(Tensor out, const Tensor in1, const Tensor in2, const Tensor t) {
  device::setDevice(out->getDevice());
  size_t length = out->shape().elements();
  gHighwayForward<<<gridFor(length, 256), 256, 0, cudaStreamOfEngine()>>>(out->data(), in1->data(), in2->data(), t->data(), length);
  CUDA_LAUNCH_CHECK();
}

void HighwayBackward(Tensor out1, Tensor out2, Tensor outt, const Tensor in1, const Tensor in2, const Tensor t, const Tensor adj) {
  device::setDevice(out1->getDevice());
  size_t length = out1->shape().elements();
  gHighwayBackward<<<gridFor(length, 256), 256, 0, cudaStreamOfEngine()>>>(
      out1->data(), out2->data(), outt->data(), in1->data(), in2->data(), t->data(), adj->data(), length);
  CUDA_LAUNCH_CHECK();
}

// =============================================================================
// Data movement: transpose / concat / rows / shift
// reference: tensor_operators.cu:35-200, 656-746, 1676-1707
// =============================================================================
namespace {

struct Perm {
  int p[4];
};

// generic <=4-D permutation; one element per thread
__global__ void gTransposeGeneric(float* __restrict__ out, const float* __restrict__ in, Shape4 os, Shape4 is, Perm permute) {
  int length = os.elements();
  for(int index = blockIdx.x * blockDim.x + threadIdx.x; index < length; index += gridDim.x * blockDim.x) {
    int oDims[4], pDims[4];
    os.dims(index, oDims);
#pragma unroll
    for(int i = 0; i < 4; ++i)
      pDims[permute.p[i]] = oDims[i];
    out[index] = in[is.index(pDims)];
  }
}

// permutations that keep the last axis (e.g. {0,2,1,3}: head split/join,
// time<->batch): whole rows move, four floats per thread
__global__ void gTransposeRows4(float4* __restrict__ out, const float4* __restrict__ in, Shape4 os, Shape4 is, Perm permute, int cols4, __nv_bfloat16* __restrict__ outShadow) {
  pdlEnter();
  long long items = (long long)os.d[0] * os.d[1] * os.d[2] * cols4;
  for(long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += (long long)gridDim.x * blockDim.x) {
    int c = (int)(w % cols4);
    int row = (int)(w / cols4);
    int oDims[4], pDims[4];
    oDims[2] = row % os.d[2];
    int t = row / os.d[2];
    oDims[1] = t % os.d[1];
    oDims[0] = t / os.d[1];
    oDims[3] = 0;
#pragma unroll
    for(int i = 0; i < 4; ++i)
      pDims[permute.p[i]] = oDims[i];
    size_t src = ((size_t)pDims[0] * is.st[0] + (size_t)pDims[1] * is.st[1] + (size_t)pDims[2] * is.st[2]) / 4;
    const float4 v = in[src + c];
    out[w] = v;
    shadow::store4(outShadow, (size_t)w << 2, v);  // bf16 copy when the transposed tensor feeds a product (BF16S)
  }
}

// swap of the two innermost axes through a padded shared-memory tile
__global__ void gTransposeLast2(float* __restrict__ out, const float* __restrict__ in, int batch, int rows, int cols) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  const float* src = in + (size_t)b * rows * cols;
  float* dst = out + (size_t)b * rows * cols;
  int x = blockIdx.x * 32 + threadIdx.x;
  int y0 = blockIdx.y * 32;
  for(int i = threadIdx.y; i < 32; i += blockDim.y) {
    int y = y0 + i;
    if(x < cols && y < rows)
      tile[i][threadIdx.x] = src[(size_t)y * cols + x];
  }
  __syncthreads();
  int ox = blockIdx.y * 32 + threadIdx.x;  // along rows of the input
  int oy0 = blockIdx.x * 32;               // along cols of the input
  for(int i = threadIdx.y; i < 32; i += blockDim.y) {
    int oy = oy0 + i;
    if(ox < rows && oy < cols)
      dst[(size_t)oy * rows + ox] = tile[threadIdx.x][i];
  }
}

// out viewed as [rows][outWidth]; copies (or reads back) a [rows][width] block at column `offset`
template <bool TO_WIDE>
__global__ void gCopyBlock(float* __restrict__ wide, float* __restrict__ narrow, int rows, int width, int outWidth, int offset) {
  long long items = (long long)rows * width;
  for(long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += (long long)gridDim.x * blockDim.x) {
    int r = (int)(w / width);
    int c = (int)(w - (long long)r * width);
    size_t wi = (size_t)r * outWidth + offset + c;
    if(TO_WIDE)
      wide[wi] = narrow[w];
    else
      narrow[w] = wide[wi];
  }
}
template <bool TO_WIDE>
__global__ void gCopyBlock4(float4* __restrict__ wide, float4* __restrict__ narrow, int rows, int width4, int outWidth4, int offset4) {
  long long items = (long long)rows * width4;
  for(long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += (long long)gridDim.x * blockDim.x) {
    int r = (int)(w / width4);
    int c = (int)(w - (long long)r * width4);
    size_t wi = (size_t)r * outWidth4 + offset4 + c;
    if(TO_WIDE)
      wide[wi] = narrow[w];
    else
      narrow[w] = wide[wi];
  }
}

template <bool TO_WIDE>
void copyBlock(float* wide, float* narrow, int rows, int width, int outWidth, int offset) {
  auto st = cudaStreamOfEngine();
  bool vec = width % 4 == 0 && outWidth % 4 == 0 && offset % 4 == 0 && (((uintptr_t)wide | (uintptr_t)narrow) & 15) == 0;
  if(vec)
    gCopyBlock4<TO_WIDE><<<gridFor((size_t)rows * width / 4, 256), 256, 0, st>>>((float4*)wide, (float4*)narrow, rows, width / 4, outWidth / 4, offset / 4);
  else
    gCopyBlock<TO_WIDE><<<gridFor((size_t)rows * width, 256), 256, 0, st>>>(wide, narrow, rows, width, outWidth, offset);
  CUDA_LAUNCH_CHECK();
}

template <bool SCATTER>
__global__ void gRows(float* __restrict__ out, const float* __restrict__ in, int cols, const int* __restrict__ idx, int rows) {
  pdlEnter();
  // one warp per row; SCATTER: out[idx[j]] += in[j] (atomic, rows may repeat)
  int warpsPerBlock = blockDim.x >> 5;
  int lane = threadIdx.x & 31;
  for(int j = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); j < rows; j += gridDim.x * warpsPerBlock) {
    int r = idx[j];
    if(SCATTER) {
      float* o = out + (size_t)r * cols;
      const float* i = in + (size_t)j * cols;
      for(int c = lane; c < cols; c += 32)
        atomicAdd(o + c, i[c]);
    } else {
      float* o = out + (size_t)j * cols;
      const float* i = in + (size_t)r * cols;
      if((cols & 3) == 0 && ((((uintptr_t)o) | ((uintptr_t)i)) & 15) == 0) {
        for(int c = lane; c < (cols >> 2); c += 32)
          reinterpret_cast<float4*>(o)[c] = reinterpret_cast<const float4*>(i)[c];
      } else {
        for(int c = lane; c < cols; c += 32)
          o[c] = i[c];
      }
    }
  }
}

__global__ void gShift(float* __restrict__ out, const float* __restrict__ in, int length, int offset) {
  pdlEnter();
  for(int index = blockIdx.x * blockDim.x + threadIdx.x; index < length; index += gridDim.x * blockDim.x) {
    if(index - offset < 0 || index - offset >= length)
      out[index] = 0;
    else
      out[index] = in[index - offset];
  }
}

__global__ void gCopyCols(float* out, const float* in, size_t rows, size_t colsIn, const int* idx, size_t colsOut, bool paste) {
  size_t n = rows * colsOut;
  for(size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (size_t)gridDim.x * blockDim.x) {
    size_t j = w / colsOut, i = w % colsOut;
    if(paste)
      out[j * colsIn + idx[i]] = in[j * colsOut + i];  // `out` is the wide tensor
    else
      out[j * colsOut + i] = in[j * colsIn + idx[i]];
  }
}

__global__ void gSelect(float* out, Shape4 os, const float* in, Shape4 is, int axis, const int* idx, bool insert) {
  // select: out[.., i, ..] = in[.., idx[i], ..];  insert: out[.., idx[i], ..] += in[.., i, ..]
  Shape4 iter = insert ? is : os;
  int length = iter.elements();
  for(int index = blockIdx.x * blockDim.x + threadIdx.x; index < length; index += gridDim.x * blockDim.x) {
    int dims[4];
    iter.dims(index, dims);
    dims[axis] = idx[dims[axis]];
    if(insert)
      atomicAdd(out + os.index(dims), in[index]);
    else
      out[index] = in[is.index(dims)];
  }
}

// temporary device copy of a host index vector (API-compat paths only)
struct TempIndices {
  int* d;
  TempIndices(const std::vector<size_t>& v) {
    std::vector<int> h(v.begin(), v.end());
    d = (int*)device::mallocDevice(h.size() * sizeof(int));
    device::copyH2DBlocking(d, h.data(), h.size() * sizeof(int));
  }
  ~TempIndices() {
    device::synchronize();
    device::freeDevice(d);
  }
};

}  // namespace

void TransposeND(Tensor out, Tensor in, const std::vector<int>& vAxis) {
  device::setDevice(out->getDevice());
  out->takeLazyZero();  // assigns every element
  Shape4 os(out->shape()), is(in->shape());
  Perm perm;
  int diff = 4 - (int)vAxis.size();
  for(int i = 0; i < 4; ++i)
    perm.p[i] = i < diff ? i : vAxis[i - diff] + diff;
  int length = os.elements();
  if(length == 0)
    return;
  auto st = cudaStreamOfEngine();

  bool keepsLast = perm.p[3] == 3;
  bool swapsLast2 = perm.p[0] == 0 && perm.p[1] == 1 && perm.p[2] == 3 && perm.p[3] == 2;
  if(keepsLast && os.d[3] % 4 == 0 && ((((uintptr_t)out->data()) | ((uintptr_t)in->data())) & 15) == 0) {
    int cols4 = os.d[3] / 4;
    launchPdl(gTransposeRows4, dim3(gridFor((size_t)length / 4, 256)), dim3(256), 0, st, (float4*)out->data(), (const float4*)in->data(), os, is, perm, cols4, shadow::produce(out));
  } else if(swapsLast2) {
    int batch = is.d[0] * is.d[1], rows = is.d[2], cols = is.d[3];
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
    gTransposeLast2<<<grid, dim3(32, 8), 0, st>>>(out->data(), in->data(), batch, rows, cols);
  } else {
    gTransposeGeneric<<<gridFor(length, 256), 256, 0, st>>>(out->data(), in->data(), os, is, perm);
  }
  CUDA_LAUNCH_CHECK();
}

namespace {
// All blocks of a concatenation in ONE launch: the pointer table travels in the kernel
// parameters (the reference issues one cudaMemcpy / kernel per input plus a stream sync,
// tensor_operators.cu:35-162; a recurrent layer concatenates one state per time step).
constexpr int kConcatMax = 96;
struct ConcatTable {
  float* narrow[kConcatMax];
  int width[kConcatMax];   // elements (or float4) per row of input i
  int offset[kConcatMax];  // its first column inside the wide row
};

// grid = (chunks, inputs); wide viewed as [rows][outWidth]
__device__ __forceinline__ void concatShadow(__nv_bfloat16* sh, size_t at, float4 v) { shadow::store4(sh, 4 * at, v); }
__device__ __forceinline__ void concatShadow(__nv_bfloat16*, size_t, float) {}  // (scalar path: the consumer converts)

template <bool TO_WIDE, typename T>
__global__ void __launch_bounds__(256) gCopyBlocks(T* __restrict__ wide, const __grid_constant__ ConcatTable tab, int rows, int outWidth, __nv_bfloat16* __restrict__ wideShadow) {
  const int i = blockIdx.y;
  T* __restrict__ narrow = (T*)tab.narrow[i];
  const int width = tab.width[i], offset = tab.offset[i];
  const long long items = (long long)rows * width;
  for(long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += (long long)gridDim.x * blockDim.x) {
    int r = (int)(w / width);
    int c = (int)(w - (long long)r * width);
    size_t wi = (size_t)r * outWidth + offset + c;
    if(TO_WIDE) {
      const T v = narrow[w];
      wide[wi] = v;
      if(wideShadow)  // the concatenation feeds a product ([U | Ux] of a recurrent cell, the states of a layer): leave its bf16 copy
        concatShadow(wideShadow, wi, v);
    } else {
      narrow[w] = wide[wi];
    }
  }
}

template <bool TO_WIDE>
void copyBlocks(float* wide, const std::vector<float*>& narrow, const std::vector<int>& widths, int rows, int outWidth, __nv_bfloat16* wideShadow = nullptr) {
  auto st = cudaStreamOfEngine();
  size_t n = narrow.size();
  int offset = 0;
  for(size_t first = 0; first < n; first += kConcatMax) {
    size_t count = std::min<size_t>(kConcatMax, n - first);
    if(count == 1 && !wideShadow) {  // a lone block keeps the plain kernel
      copyBlock<TO_WIDE>(wide, narrow[first], rows, widths[first], outWidth, offset);
      offset += widths[first];
      continue;
    }
    ConcatTable tab;
    bool vec = outWidth % 4 == 0 && (((uintptr_t)wide) & 15) == 0;
    int maxWidth = 0;
    for(size_t k = 0; k < count; ++k) {
      tab.narrow[k] = narrow[first + k];
      tab.width[k] = widths[first + k];
      tab.offset[k] = offset;
      vec = vec && widths[first + k] % 4 == 0 && offset % 4 == 0 && (((uintptr_t)narrow[first + k]) & 15) == 0;
      maxWidth = std::max(maxWidth, widths[first + k]);
      offset += widths[first + k];
    }
    if(vec)
      for(size_t k = 0; k < count; ++k) {
        tab.width[k] /= 4;
        tab.offset[k] /= 4;
      }
    size_t items = (size_t)rows * (vec ? maxWidth / 4 : maxWidth);
    // enough blocks over all inputs together to fill the machine, at least one per input
    int perInput = (int)std::max<size_t>(1, std::min<size_t>((items + 255) / 256, (size_t)(kNumSMs * 8 + count - 1) / count));
    dim3 grid(perInput, (unsigned)count);
    ABORT_IF(wideShadow && !vec, "concatenation: a bf16 copy was promised but the blocks are not 16-byte aligned");
    if(vec)
      gCopyBlocks<TO_WIDE, float4><<<grid, 256, 0, st>>>((float4*)wide, tab, rows, outWidth / 4, wideShadow);
    else
      gCopyBlocks<TO_WIDE, float><<<grid, 256, 0, st>>>(wide, tab, rows, outWidth, nullptr);
    CUDA_LAUNCH_CHECK();
  }
}
}  // namespace

void Concatenate(Tensor out, const std::vector<Tensor>& inputs, int ax) {
  device::setDevice(out->getDevice());
  // rows = product of dims before `ax`; each input contributes a contiguous
  // block of (elements / rows) per row.  One launch for (up to 96) inputs, no sync.
  int rows = 1;
  for(int i = 0; i < ax; ++i)
    rows *= out->shape()[i];
  int outWidth = out->shape().elements() / rows;
  std::vector<float*> ptrs;
  std::vector<int> widths;
  // every block 16-byte aligned: the copy can leave the bf16 copy a consuming product wants (BF16S mode)
  bool vec = outWidth % 4 == 0 && (((uintptr_t)out->data()) & 15) == 0;
  int offset = 0;
  for(auto in : inputs) {
    ptrs.push_back(in->data());
    widths.push_back(in->shape().elements() / rows);
    vec = vec && widths.back() % 4 == 0 && offset % 4 == 0 && (((uintptr_t)ptrs.back()) & 15) == 0;
    offset += widths.back();
  }
  out->takeLazyZero();  // every element is written
  copyBlocks<true>(out->data(), ptrs, widths, rows, outWidth, vec ? shadow::produce(out) : nullptr);
}

void Deconcatenate(std::vector<Tensor>& outputs, const Tensor in, int ax) {
  device::setDevice(in->getDevice());
  int rows = 1;
  for(int i = 0; i < ax; ++i)
    rows *= in->shape()[i];
  int inWidth = in->shape().elements() / rows;
  std::vector<float*> ptrs;
  std::vector<int> widths;
  for(auto out : outputs) {
    out->takeLazyZero();  // ASSIGNS
    ptrs.push_back(out->rawData());
    widths.push_back(out->shape().elements() / rows);
  }
  copyBlocks<false>(in->data(), ptrs, widths, rows, inWidth);
}

void CopyRows(Tensor out, const Tensor in, const int* deviceIndices, size_t n) {
  device::setDevice(out->getDevice());
  int cols = in->shape().back();
  launchPdl(gRows<false>, dim3(gridFor(n * 32, 256)), dim3(256), 0, cudaStreamOfEngine(), out->data(), (const float*)in->data(), cols, deviceIndices, (int)n);
  CUDA_LAUNCH_CHECK();
}

void PasteRows(Tensor out, const Tensor in, const int* deviceIndices, size_t n) {
  device::setDevice(out->getDevice());
  int cols = in->shape().back();
  launchPdl(gRows<true>, dim3(gridFor(n * 32, 256)), dim3(256), 0, cudaStreamOfEngine(), out->data(), (const float*)in->data(), cols, deviceIndices, (int)n);
  CUDA_LAUNCH_CHECK();
}

void CopyRows(Tensor out, const Tensor in, const std::vector<size_t>& indices) {
  TempIndices idx(indices);
  CopyRows(out, in, idx.d, indices.size());
}
void PasteRows(Tensor out, const Tensor in, const std::vector<size_t>& indices) {
  TempIndices idx(indices);
  PasteRows(out, in, idx.d, indices.size());
}

void CopyCols(Tensor out, const Tensor in, const std::vector<size_t>& indices) {
  device::setDevice(out->getDevice());
  TempIndices idx(indices);
  size_t colsIn = in->shape().back(), colsOut = indices.size();
  size_t rows = in->shape().elements() / colsIn;
  gCopyCols<<<gridFor(rows * colsOut, 256), 256, 0, cudaStreamOfEngine()>>>(out->data(), in->data(), rows, colsIn, idx.d, colsOut, false);
  CUDA_LAUNCH_CHECK();
}
void PasteCols(Tensor out, const Tensor in, const std::vector<size_t>& indices) {
  device::setDevice(out->getDevice());
  TempIndices idx(indices);
  size_t colsWide = out->shape().back(), colsNarrow = indices.size();
  size_t rows = out->shape().elements() / colsWide;
  gCopyCols<<<gridFor(rows * colsNarrow, 256), 256, 0, cudaStreamOfEngine()>>>(out->data(), in->data(), rows, colsWide, idx.d, colsNarrow, true);
  CUDA_LAUNCH_CHECK();
}

void Select(Ptr<Allocator>, Tensor out, Tensor in, int axis, const std::vector<size_t>& indices) {
  device::setDevice(out->getDevice());
  TempIndices idx(indices);
  Shape4 os(out->shape()), is(in->shape());
  int ax = axis + 4 - (int)out->shape().size();
  gSelect<<<gridFor(os.elements(), 256), 256, 0, cudaStreamOfEngine()>>>(out->data(), os, in->data(), is, ax, idx.d, false);
  CUDA_LAUNCH_CHECK();
}
void Insert(Ptr<Allocator>, Tensor out, Tensor in, int axis, const std::vector<size_t>& indices) {
  device::setDevice(out->getDevice());
  TempIndices idx(indices);
  Shape4 os(out->shape()), is(in->shape());
  int ax = axis + 4 - (int)out->shape().size();
  gSelect<<<gridFor(is.elements(), 256), 256, 0, cudaStreamOfEngine()>>>(out->data(), os, in->data(), is, ax, idx.d, true);
  CUDA_LAUNCH_CHECK();
}

void Shift(Tensor out, Tensor in, Shape shift, bool invert) {
  ABORT_IF(in->shape().size() != shift.size(), "bad dimensions");
  device::setDevice(out->getDevice());
  int offset = 0;
  for(int i = 0; i < (int)shift.size(); ++i)
    offset += in->shape().stride(i) * shift[i];
  if(invert)
    offset = -offset;
  out->takeLazyZero();  // assigns every element
  int length = out->shape().elements();
  launchPdl(gShift, dim3(gridFor(length, 256)), dim3(256), 0, cudaStreamOfEngine(), out->data(), (const float*)in->data(), length, offset);
  CUDA_LAUNCH_CHECK();
}

// =============================================================================
// Norms, dropout, fused optimizer steps
// reference: tensor_operators.cu:1286-1305 (L2Norm), kernels/dropout.cu,
//            optimizers/optimizers.cu:7-73, optimizers/clippers.cu:12-17
// =============================================================================
namespace {

__global__ void __launch_bounds__(256) gSumSquares(float* __restrict__ out, const float* __restrict__ in, size_t n) {
  pdlEnter();
  __shared__ float smem[32];
  float acc = 0.f;
  size_t n4 = n >> 2;
  const float4* p4 = reinterpret_cast<const float4*>(in);
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 q = p4[i];
    acc += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  }
  for(size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += in[i] * in[i];
  acc = blockSum(acc, smem);
  if(threadIdx.x == 0)
    atomicAdd(out, acc);
}

__device__ __forceinline__ float clipScale(float gradScale, float clipNorm, const float* normSq) {
  // Norm::clip: if(|g| >= c) g *= c / |g|, on the already scaled gradient
  float scale = gradScale;
  if(clipNorm > 0.f && normSq) {
    float norm = sqrtf(*normSq) * gradScale;
    if(norm >= clipNorm)
      scale *= clipNorm / norm;
  }
  return scale;
}

__global__ void __launch_bounds__(256) gAdam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n, AdamArgs a, const float* __restrict__ normSq, PeerStores peers) {
  pdlEnter();
  float scale = clipScale(a.gradScale, a.clipNorm, normSq);
  size_t n4 = n >> 2;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x;
    float* G = &gg.x;
    float* M = &mm.x;
    float* V = &vv.x;
#pragma unroll
    for(int e = 0; e < 4; ++e) {
      float gi = G[e] * scale;
      M[e] = (a.beta1 * M[e]) + ((1 - a.beta1) * gi);
      V[e] = (a.beta2 * V[e]) + ((1 - a.beta2) * (gi * gi));
      P[e] = P[e] - a.eta * (M[e] / a.denom1) / (sqrtf(V[e] / a.denom2) + a.eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    shadow::store4((__nv_bfloat16*)a.shadow, i << 2, pp);  // bf16 copy for the tensor-core products (BF16S mode)
    // all-gather by peer stores: the owner writes the new values into every replica
    for(int r = 0; r < peers.nranks; ++r)
      if(r != peers.self)
        reinterpret_cast<float4*>(reinterpret_cast<float*>(peers.params.ptr[r]) + peers.offset)[i] = pp;
  }
  for(size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * scale;
    m[i] = (a.beta1 * m[i]) + ((1 - a.beta1) * gi);
    v[i] = (a.beta2 * v[i]) + ((1 - a.beta2) * (gi * gi));
    p[i] = p[i] - a.eta * (m[i] / a.denom1) / (sqrtf(v[i] / a.denom2) + a.eps);
    shadow::store1((__nv_bfloat16*)a.shadow, i, p[i]);
    for(int r = 0; r < peers.nranks; ++r)
      if(r != peers.self)
        (reinterpret_cast<float*>(peers.params.ptr[r]) + peers.offset)[i] = p[i];
  }
}

__global__ void gSgd(float* p, const float* g, size_t n, float eta, float gradScale, float clipNorm, const float* normSq) {
  float scale = clipScale(gradScale, clipNorm, normSq);
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] -= eta * (g[i] * scale);
}

__global__ void gAdagrad(float* p, const float* g, float* gt, size_t n, float eta, float eps, float gradScale, float clipNorm, const float* normSq) {
  float scale = clipScale(gradScale, clipNorm, normSq);
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * scale;
    float acc = gt[i] + gi * gi;
    gt[i] = acc;
    p[i] -= (eta / (sqrtf(acc) + eps)) * gi;
  }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__global__ void gDropoutEpochBump(uint64_t* epoch) {
  if(threadIdx.x == 0 && blockIdx.x == 0)
    *epoch += 1;
}
__global__ void gDropout(float* mask, size_t n, float p, float scale, uint64_t seed, const uint64_t* epoch) {
  if(epoch)
    seed = mix64(seed ^ (*epoch * 0xD6E8FEB86659FD93ULL));
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ULL * (i + 1));
    float u = (float)(h >> 40) * (1.0f / 16777216.0f);  // [0,1)
    mask[i] = (u >= p) ? scale : 0.f;
  }
}

bool all16(std::initializer_list<const void*> ps) {
  for(auto p : ps)
    if(((uintptr_t)p & 15) != 0)
      return false;
  return true;
}

}  // namespace

void SumSquares(Tensor outScalar, Tensor in) {
  device::setDevice(in->getDevice());
  size_t n = in->size();
  device::zero(outScalar->data(), sizeof(float));
  ABORT_IF(((uintptr_t)in->data() & 15) != 0, "SumSquares expects a 16-byte aligned tensor");
  int grid = std::max(1, std::min((int)((n / 4 + 255) / 256), kNumSMs * 8));
  launchPdl(gSumSquares, dim3(grid), dim3(256), 0, cudaStreamOfEngine(), outScalar->data(), (const float*)in->data(), n);
  CUDA_LAUNCH_CHECK();
}

float L2Norm(Tensor in) {
  device::setDevice(in->getDevice());
  float* d = (float*)device::mallocDevice(256);
  auto mem = New<MemoryPiece>((uint8_t*)d, sizeof(float));
  Tensor s(new TensorBase(mem, Shape{1, 1}, in->getDevice()));
  SumSquares(s, in);
  float v = s->scalar();
  device::freeDevice(d);
  return sqrtf(v);
}

void Dropout(Tensor mask, float dropProb, uint64_t seed, const uint64_t* epoch) {
  device::setDevice(mask->getDevice());
  size_t n = mask->size();
  gDropout<<<gridFor(n, 256), 256, 0, cudaStreamOfEngine()>>>(mask->data(), n, dropProb, 1.f / (1.f - dropProb), seed, epoch);
  CUDA_LAUNCH_CHECK();
}
void DropoutEpochBump(uint64_t* epoch) {
  gDropoutEpochBump<<<1, 32, 0, cudaStreamOfEngine()>>>(epoch);
  CUDA_LAUNCH_CHECK();
}

void AdamUpdate(Tensor params, Tensor grads, Tensor mt, Tensor vt, const AdamArgs& args, Tensor normSq, const PeerStores* peers) {
  device::setDevice(params->getDevice());
  size_t n = params->size();
  ABORT_IF(grads->size() != n || mt->size() != n || vt->size() != n, "AdamUpdate: size mismatch");
  ABORT_IF(!all16({params->data(), grads->data(), mt->data(), vt->data()}), "AdamUpdate expects 16-byte aligned tensors");
  int grid = std::max(1, std::min((int)((n / 4 + 255) / 256), kNumSMs * 8));
  launchPdl(gAdam, dim3(grid), dim3(256), 0, cudaStreamOfEngine(), params->data(), (const float*)grads->data(), mt->data(), vt->data(), n, args, (const float*)(normSq ? normSq->data() : nullptr), peers ? *peers : PeerStores());
  CUDA_LAUNCH_CHECK();
}

void SgdUpdate(Tensor params, Tensor grads, float eta, float gradScale, float clipNorm, Tensor normSq) {
  device::setDevice(params->getDevice());
  size_t n = params->size();
  gSgd<<<gridFor(n, 256), 256, 0, cudaStreamOfEngine()>>>(params->data(), grads->data(), n, eta, gradScale, clipNorm, normSq ? normSq->data() : nullptr);
  CUDA_LAUNCH_CHECK();
}

void AdagradUpdate(Tensor params, Tensor grads, Tensor gt, float eta, float eps, float gradScale, float clipNorm, Tensor normSq) {
  device::setDevice(params->getDevice());
  size_t n = params->size();
  gAdagrad<<<gridFor(n, 256), 256, 0, cudaStreamOfEngine()>>>(params->data(), grads->data(), gt->data(), n, eta, eps, gradScale, clipNorm, normSq ? normSq->data() : nullptr);
  CUDA_LAUNCH_CHECK();
}

}  // namespace marian
