// Element / Add / Reduce: the fused element-wise kernel family (CUDA, sm_100a).
//
// Call syntax and dispatch rules of the reference's templates
// (src/kernels/tensor_operators.h:26-274):
//   Element(f, out, ins...)          out[i] = f(out[i], ins[bcast(i)]...)
//   Add(f, [scale,] out, ins...)     out += scale * reduce_or_broadcast(f(ins...))
//      (1) full.back()!=1 && out.back()==1 : last-axis reduction   (gAddReduce there)
//      (2) out.shape()==full               : element-wise accumulate (gAddEqual)
//      (3) otherwise                       : generic reduction       (gAddGeneric)
//   Reduce = out->set(0); Add(...)
//
// Kernel design (all HBM-bound):
//  * one kernel template for Element and Add-case-(2): each thread owns FOUR
//    consecutive elements of a row; the broadcast row base of every operand is
//    resolved once per thread (the reference does a 4-D div/mod per element per
//    operand), contiguous operands are read with 128-bit loads, operands whose
//    last dim is 1 with one scalar load.  Operands the functor never reads
//    (`_1 = _2 + _3` does not read out) are not loaded (functional::Reads).
//  * case (1): one warp per row, lanes stride the row, shuffle reduction
//    (the reference: 512-thread blocks with a shared-memory tree per row).
//  * case (3): one thread per OUTPUT element so that consecutive threads read
//    consecutive addresses of every operand, the reduced sub-space is split over
//    gridDim.y slices that combine with atomicAdd (the reference runs the whole
//    reduction serially in one thread per output: the bias gradient of a
//    [3200,512] adjoint ran on 512 threads).
// Grids are sized in multiples of the 148 SMs.
#pragma once

#include <cuda_runtime.h>

#include "common/shape.h"
#include "functional/functional.h"
#include "kernels/cuda_helpers.h"
#include "kernels/shadow.h"
#include "tensors/tensor.h"

namespace marian {
namespace ew {

template <int K>
struct Operands {
  const float* p[K];
  int rb[K][3];  // broadcast strides of dims 0..2 (row base)
  int cs[K];     // stride of the last dim: 1, or 0 when broadcast along it
};

struct RowGeom {
  int rows, cols;  // rows = product of dims 0..2 of the iteration shape
  int d1, d2;      // extents of dims 1 and 2 (row decode)
};

// MODE 0: Element (operand 0 is `out`, out = f(...));  MODE 1: Add, out += scale f(ins);
// MODE 2: Add into a lazily-zero output, out = scale f(ins) (first writer assigns)
// FLAT: every operand has exactly the iteration shape (one long row, g.rows == 1): no index
// decode at all, the kernel is a pure 128-bit stream.
template <int K, int MODE, bool VEC, class Functor, bool FLAT = false>
__global__ void __launch_bounds__(256) gElementwise(Functor f, float* __restrict__ out, Operands<K> ops, RowGeom g, float scale, __nv_bfloat16* __restrict__ outShadow) {
  pdlEnter();
  const unsigned cpr = (unsigned)(g.cols + 3) >> 2;  // 4-element chunks per row
  const unsigned items = (unsigned)g.rows * cpr;     // < 2^31 (checked by the launchers)
  for(unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < items; w += gridDim.x * blockDim.x) {
    int row = FLAT ? 0 : (int)(w / cpr);
    int c = FLAT ? (int)(w << 2) : (int)(w - (unsigned)row * cpr) << 2;
    int o2 = FLAT ? 0 : row % g.d2;
    int t = FLAT ? 0 : row / g.d2;
    int o1 = FLAT ? 0 : t % g.d1;
    int o0 = FLAT ? 0 : t / g.d1;

    float v[K][4];
#pragma unroll
    for(int k = 0; k < K; ++k) {
      // Element mode: operand 0 is `out`; skip the load if the functor ignores it
      if(MODE == 0 && k == 0 && !functional::Reads<Functor, 1>::value) {
        v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0.f;
        continue;
      }
      const float* base = FLAT ? ops.p[k] : ops.p[k] + (size_t)o0 * ops.rb[k][0] + (size_t)o1 * ops.rb[k][1] + (size_t)o2 * ops.rb[k][2];
      if(!FLAT && ops.cs[k] == 0) {
        float s = __ldg(base);
        v[k][0] = v[k][1] = v[k][2] = v[k][3] = s;
      } else if(VEC) {
        float4 q = *reinterpret_cast<const float4*>(base + c);
        v[k][0] = q.x;
        v[k][1] = q.y;
        v[k][2] = q.z;
        v[k][3] = q.w;
      } else {
#pragma unroll
        for(int e = 0; e < 4; ++e)
          v[k][e] = (c + e < g.cols) ? base[c + e] : 0.f;
      }
    }

    float r[4];
#pragma unroll
    for(int e = 0; e < 4; ++e) {
      float a[K];
#pragma unroll
      for(int k = 0; k < K; ++k)
        a[k] = v[k][e];
      r[e] = f(a);
    }

    float* o = out + (size_t)row * g.cols + c;
    if(VEC) {
      float4 q;
      if(MODE == 1) {
        q = *reinterpret_cast<float4*>(o);
        q.x += r[0] * scale;
        q.y += r[1] * scale;
        q.z += r[2] * scale;
        q.w += r[3] * scale;
      } else if(MODE == 2) {
        q = make_float4(r[0] * scale, r[1] * scale, r[2] * scale, r[3] * scale);
      } else {
        q = make_float4(r[0], r[1], r[2], r[3]);
      }
      if(out)  // (null: shadow-only output, see shadow::fp32Target)
        *reinterpret_cast<float4*>(o) = q;
      shadow::store4(outShadow, (size_t)row * g.cols + c, q);  // bf16 copy when `out` feeds a tensor-core product (BF16S)
    } else {
#pragma unroll
      for(int e = 0; e < 4; ++e)
        if(c + e < g.cols) {
          if(MODE == 1)
            o[e] += r[e] * scale;
          else if(MODE == 2)
            o[e] = r[e] * scale;
          else
            o[e] = r[e];
        }
    }
  }
}

// case (1): out[row] += scale * sum_c f(ins[row, c]); one warp per row
template <int K, class Functor>
__global__ void __launch_bounds__(256) gAddReduceRows(Functor f, float* __restrict__ out, Operands<K> ops, RowGeom g, float scale, int assign) {
  pdlEnter();
  int warpsPerBlock = blockDim.x >> 5;
  int lane = threadIdx.x & 31;
  for(int row = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); row < g.rows; row += gridDim.x * warpsPerBlock) {
    int o2 = row % g.d2;
    int t = row / g.d2;
    int o1 = t % g.d1;
    int o0 = t / g.d1;
    const float* base[K];
#pragma unroll
    for(int k = 0; k < K; ++k)
      base[k] = ops.p[k] + (size_t)o0 * ops.rb[k][0] + (size_t)o1 * ops.rb[k][1] + (size_t)o2 * ops.rb[k][2];
    float acc = 0.f;
    for(int c = lane; c < g.cols; c += 32) {
      float a[K];
#pragma unroll
      for(int k = 0; k < K; ++k)
        a[k] = base[k][(size_t)c * ops.cs[k]];
      acc += f(a);
    }
    acc = warpSum(acc);
    if(lane == 0)
      out[row] = assign ? acc * scale : out[row] + acc * scale;
  }
}

// case (3), column form: out[1,1,1,C] += scale * sum over all rows of f(ins[row, c]) - bias,
// gamma and beta gradients.  blockDim = (32, 8): a warp covers 128 consecutive columns with
// 128-bit loads, the 8 warps stride the rows of a slice, partial sums meet in shared memory
// and leave through ONE atomicAdd per column and slice.
template <int K, class Functor>
__global__ void __launch_bounds__(256) gAddColumns(Functor f, float* __restrict__ out, Operands<K> ops, RowGeom g, float scale, int rowsPerSlice, int assign, int keep2,
                                                   __nv_bfloat16* __restrict__ outShadow) {
  pdlEnter();
  __shared__ float red[8][32][5];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  if(keep2 && outShadow)
    outShadow += (size_t)blockIdx.z * g.cols;
  const int r0 = blockIdx.y * rowsPerSlice;
  const int r1 = min(g.rows, r0 + rowsPerSlice);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if(keep2)  // out[.., o2, C]: the third axis is kept (blockIdx.z), rows run over the two leading axes only
    out += (size_t)blockIdx.z * g.cols;
  if(c < g.cols) {
    for(int row = r0 + threadIdx.y; row < r1; row += 8) {
      int o2, o1, o0;
      if(keep2) {
        o2 = blockIdx.z;
        o1 = row % g.d1;
        o0 = row / g.d1;
      } else {
        o2 = row % g.d2;
        int t = row / g.d2;
        o1 = t % g.d1;
        o0 = t / g.d1;
      }
      float v[K][4];
#pragma unroll
      for(int k = 0; k < K; ++k) {
        const float* base = ops.p[k] + (size_t)o0 * ops.rb[k][0] + (size_t)o1 * ops.rb[k][1] + (size_t)o2 * ops.rb[k][2];
        if(ops.cs[k] == 0) {
          float sv = __ldg(base);
          v[k][0] = v[k][1] = v[k][2] = v[k][3] = sv;
        } else {
          float4 q = *reinterpret_cast<const float4*>(base + c);
          v[k][0] = q.x;
          v[k][1] = q.y;
          v[k][2] = q.z;
          v[k][3] = q.w;
        }
      }
#pragma unroll
      for(int e = 0; e < 4; ++e) {
        float a[K];
#pragma unroll
        for(int k = 0; k < K; ++k)
          a[k] = v[k][e];
        acc[e] += f(a);
      }
    }
  }
#pragma unroll
  for(int e = 0; e < 4; ++e)
    red[threadIdx.y][threadIdx.x][e] = acc[e];
  __syncthreads();
  if(threadIdx.y == 0 && c < g.cols) {
    float sum[4];
#pragma unroll
    for(int e = 0; e < 4; ++e) {
      sum[e] = 0.f;
#pragma unroll
      for(int y = 0; y < 8; ++y)
        sum[e] += red[y][threadIdx.x][e];
      sum[e] *= scale;
    }
    float4 v = make_float4(sum[0], sum[1], sum[2], sum[3]);
    if(assign) {
      *reinterpret_cast<float4*>(out + c) = v;
      shadow::store4(outShadow, (size_t)c, v);  // complete values: the bf16 copy a consuming product reads (BF16S mode)
    } else {
      redAdd4(out + c, v);  // out is 16-byte aligned (checked by the launcher)
    }
  }
}

struct GenericGeom {
  Shape4 out;     // output shape (4-D)
  int len[4];     // reduced extent per dim (full[i] / out[i])
  int total;      // product of len
  int chunk;      // reduced elements per slice
  int outLength;
  int atomic;     // slices > 1
  int assign;     // single slice into a lazily-zero output
};

template <int K>
struct FullOperands {
  const float* p[K];
  int bst[K][4];
};

// case (3): thread per output element, reduced sub-space split over blockIdx.y
template <int K, class Functor>
__global__ void __launch_bounds__(128) gAddGeneric(Functor f, float* __restrict__ out, FullOperands<K> ops, GenericGeom g, float scale) {
  pdlEnter();
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if(o >= g.outLength)
    return;
  int od[4];
  g.out.dims(o, od);
  int begin = blockIdx.y * g.chunk;
  int end = min(g.total, begin + g.chunk);
  if(begin >= end)
    return;

  // decode `begin` into a 4-D counter over len[], then advance with carries
  int i[4];
  {
    int r = begin;
    i[3] = r % g.len[3];
    r /= g.len[3];
    i[2] = r % g.len[2];
    r /= g.len[2];
    i[1] = r % g.len[1];
    i[0] = r / g.len[1];
  }
  size_t idx[K];
#pragma unroll
  for(int k = 0; k < K; ++k) {
    idx[k] = 0;
#pragma unroll
    for(int j = 0; j < 4; ++j)
      idx[k] += (size_t)(od[j] + i[j]) * ops.bst[k][j];
  }

  float acc = 0.f;
  for(int r = begin; r < end; ++r) {
    float a[K];
#pragma unroll
    for(int k = 0; k < K; ++k)
      a[k] = ops.p[k][idx[k]];
    acc += f(a);
    // increment the counter (dim 3 fastest) and the operand offsets
    int j = 3;
#pragma unroll
    for(int step = 0; step < 4; ++step) {
      if(j < 0)
        break;
      i[j]++;
#pragma unroll
      for(int k = 0; k < K; ++k)
        idx[k] += ops.bst[k][j];
      if(i[j] < g.len[j])
        break;
#pragma unroll
      for(int k = 0; k < K; ++k)
        idx[k] -= (size_t)g.len[j] * ops.bst[k][j];
      i[j] = 0;
      --j;
    }
  }
  if(g.atomic)
    atomicAdd(out + o, acc * scale);
  else if(g.assign)
    out[o] = acc * scale;
  else
    out[o] += acc * scale;
}

inline bool aligned16(const void* p) {
  return ((uintptr_t)p & 15) == 0;
}

// Fills row-geometry + operand strides for iterating `iter` (the shape being
// walked) with operands of (possibly broadcast) shapes.
template <int K>
inline void setupOperands(const Shape4& iter, const Tensor* ts, Operands<K>& ops, RowGeom& g, bool& vec, bool collapse) {
  if(collapse) {
    // every operand has exactly the iteration shape: walk it as one long row
    g.rows = 1;
    g.cols = iter.elements();
    g.d1 = g.d2 = 1;
    for(int k = 0; k < K; ++k) {
      ops.p[k] = ts[k]->data();
      ops.rb[k][0] = ops.rb[k][1] = ops.rb[k][2] = 0;
      ops.cs[k] = 1;
    }
  } else {
    g.cols = iter.d[3];
    g.rows = iter.d[0] * iter.d[1] * iter.d[2];
    g.d1 = iter.d[1];
    g.d2 = iter.d[2];
    for(int k = 0; k < K; ++k) {
      Shape4 s(ts[k]->shape());
      ops.p[k] = ts[k]->data();
      ops.rb[k][0] = s.bst[0];
      ops.rb[k][1] = s.bst[1];
      ops.rb[k][2] = s.bst[2];
      ops.cs[k] = s.bst[3];
    }
  }
  vec = (g.cols % 4 == 0);
  for(int k = 0; k < K; ++k)
    if(ops.cs[k] == 1 && !aligned16(ops.p[k]))
      vec = false;
}

}  // namespace ew

template <class Functor, class... Tensors>
void Element(Functor functor, Tensor out, Tensors... tensors) {
  device::setDevice(out->getDevice());
  constexpr int K = sizeof...(tensors) + 1;
  if(!functional::Reads<Functor, 1>::value)
    out->takeLazyZero();  // every element is overwritten: a pending lazy zero is moot
  Tensor ts[K] = {out, tensors...};

  Shape4 iter(out->shape());
  bool broadcast = false;
  for(int i = 1; i < K; ++i)
    broadcast = broadcast || iter != Shape4(ts[i]->shape());

  ew::Operands<K> ops;
  ew::RowGeom g;
  bool vec;
  ew::setupOperands<K>(iter, ts, ops, g, vec, !broadcast);
  if(g.rows == 0 || g.cols == 0)
    return;

  long long items = (long long)g.rows * ((g.cols + 3) / 4);
  ABORT_IF(items >= (1ll << 31), "Element: more than 2^33 elements");
  int grid = gridFor((size_t)items, 256);
  auto stream = cudaStreamOfEngine();
  // every element of `out` is (re)written: the vector kernels also leave its bf16 shadow when one is wanted
  __nv_bfloat16* osh = vec ? shadow::produce(out) : nullptr;
  // a shadow-only output (all readers are products) of a functor that does not read `out`: no fp32 stores
  float* outp = (osh && !functional::Reads<Functor, 1>::value) ? shadow::fp32Target(out, osh) : out->data();
  if(vec && !broadcast)
    launchPdl(ew::gElementwise<K, 0, true, Functor, true>, dim3(grid), dim3(256), 0, stream, functor, outp, ops, g, 1.f, osh);
  else if(vec)
    launchPdl(ew::gElementwise<K, 0, true, Functor>, dim3(grid), dim3(256), 0, stream, functor, outp, ops, g, 1.f, osh);
  else
    launchPdl(ew::gElementwise<K, 0, false, Functor>, dim3(grid), dim3(256), 0, stream, functor, out->data(), ops, g, 1.f, osh);
  CUDA_LAUNCH_CHECK();
}

template <class Functor, class... Tensors>
void Add(Functor functor, float scale, Tensor out, Tensors... tensors) {
  device::setDevice(out->getDevice());
  constexpr int K = sizeof...(Tensors);
  Tensor ts[K] = {tensors...};

  std::vector<Shape> shapes = {out->shape(), tensors->shape()...};
  Shape4 full(Shape::broadcast(shapes));
  Shape4 outS(out->shape());
  auto stream = cudaStreamOfEngine();
  if(full.elements() == 0)
    return;

  if(full.back() != 1 && outS.back() == 1) {
    // (1) reduce the last axis
    ew::Operands<K> ops;
    ew::RowGeom g;
    bool vec;
    ew::setupOperands<K>(full, ts, ops, g, vec, false);
    int warpsPerBlock = 8;
    int grid = gridFor((size_t)g.rows * 32, 256);
    (void)warpsPerBlock;
    int assign = out->takeLazyZero() ? 1 : 0;
    launchPdl(ew::gAddReduceRows<K, Functor>, dim3(grid), dim3(256), 0, stream, functor, out->data(), ops, g, scale, assign);
  } else if(outS == full) {
    // (2) element-wise accumulate
    bool broadcast = false;
    for(int i = 0; i < K; ++i)
      broadcast = broadcast || full != Shape4(ts[i]->shape());
    ew::Operands<K> ops;
    ew::RowGeom g;
    bool vec;
    ew::setupOperands<K>(full, ts, ops, g, vec, !broadcast);
    bool assign = out->takeLazyZero();
    if(!ew::aligned16(out->data()))
      vec = false;
    __nv_bfloat16* const nosh = nullptr;  // an accumulating pass is not the adjoint's only writer: no shadow
    long long items = (long long)g.rows * ((g.cols + 3) / 4);
    ABORT_IF(items >= (1ll << 31), "Add: more than 2^33 elements");
    int grid = gridFor((size_t)items, 256);
    if(vec && !broadcast) {
      if(assign)
        launchPdl(ew::gElementwise<K, 2, true, Functor, true>, dim3(grid), dim3(256), 0, stream, functor, out->data(), ops, g, scale, shadow::produce(out));
      else
        launchPdl(ew::gElementwise<K, 1, true, Functor, true>, dim3(grid), dim3(256), 0, stream, functor, out->data(), ops, g, scale, nosh);
    } else if(assign) {
      if(vec)
        launchPdl(ew::gElementwise<K, 2, true, Functor>, dim3(grid), dim3(256), 0, stream, functor, out->data(), ops, g, scale, shadow::produce(out));
      else
        launchPdl(ew::gElementwise<K, 2, false, Functor>, dim3(grid), dim3(256), 0, stream, functor, out->data(), ops, g, scale, nosh);
    } else if(vec)
      launchPdl(ew::gElementwise<K, 1, true, Functor>, dim3(grid), dim3(256), 0, stream, functor, out->data(), ops, g, scale, nosh);
    else
      launchPdl(ew::gElementwise<K, 1, false, Functor>, dim3(grid), dim3(256), 0, stream, functor, out->data(), ops, g, scale, nosh);
  } else if(outS.d[0] == 1 && outS.d[1] == 1 && outS.d[2] == 1 && outS.d[3] == full.d[3] && (full.d[3] & 3) == 0 && ew::aligned16(out->memory()->data())
            && [&] {
                 for(int k = 0; k < K; ++k) {
                   Shape4 sk(ts[k]->shape());
                   if(sk.bst[3] == 1 && (!ew::aligned16(ts[k]->data()) || (sk.d[3] & 3)))
                     return false;
                 }
                 return true;
               }()) {
    // (3a) column sums over all leading dims (bias / gamma / beta gradients)
    ew::Operands<K> ops;
    ew::RowGeom g;
    bool vec;
    ew::setupOperands<K>(full, ts, ops, g, vec, false);
    int strips = (g.cols / 4 + 31) / 32;
    int slices = std::max(1, std::min((kNumSMs * 4 + strips - 1) / strips, (g.rows + 15) / 16));
    int rowsPerSlice = (g.rows + slices - 1) / slices;
    slices = (g.rows + rowsPerSlice - 1) / rowsPerSlice;
    int assign = (slices == 1 && out->takeLazyZero()) ? 1 : 0;
    launchPdl(ew::gAddColumns<K, Functor>, dim3(strips, slices), dim3(32, 8), 0, stream, functor, out->data(), ops, g, scale, rowsPerSlice, assign, 0, (__nv_bfloat16*)nullptr);
  } else if(outS.d[0] == 1 && outS.d[1] == 1 && outS.d[2] == full.d[2] && outS.d[3] == full.d[3] && full.d[2] <= 65535 && (full.d[3] & 3) == 0 && ew::aligned16(out->memory()->data())
            && ew::aligned16(out->data()) && [&] {
                 for(int k = 0; k < K; ++k) {
                   Shape4 sk(ts[k]->shape());
                   if(sk.bst[3] == 1 && (!ew::aligned16(ts[k]->data()) || (sk.d[3] & 3)))
                     return false;
                 }
                 return true;
               }()) {
    // (3b) sums over the two leading axes, the third axis kept: out[1,1,B,C] += sum_t f(ins[t,b,c]) - the context
    // vector of the recurrent decoder's attention (scalar_product over source positions), means over time.  The
    // same strip kernel as (3a), one grid layer per kept index.
    ew::Operands<K> ops;
    ew::RowGeom g;
    bool vec;
    ew::setupOperands<K>(full, ts, ops, g, vec, false);
    g.rows = full.d[0] * full.d[1];
    int strips = (g.cols / 4 + 31) / 32;
    long layers = (long)strips * full.d[2];
    int slices = (int)std::max<long>(1, std::min<long>((kNumSMs * 4 + layers - 1) / layers, (g.rows + 15) / 16));
    int rowsPerSlice = (g.rows + slices - 1) / slices;
    slices = (g.rows + rowsPerSlice - 1) / rowsPerSlice;
    int assign = (slices == 1 && out->takeLazyZero()) ? 1 : 0;
    // (the attention context of a recurrent decoder step is the input of the next cell's product)
    __nv_bfloat16* osh = assign ? shadow::produce(out) : nullptr;
    launchPdl(ew::gAddColumns<K, Functor>, dim3(strips, slices, full.d[2]), dim3(32, 8), 0, stream, functor, out->data(), ops, g, scale, rowsPerSlice, assign, 1, osh);
  } else {
    // (3) generic reduction over the dims where out has extent 1
    ew::FullOperands<K> ops;
    for(int k = 0; k < K; ++k) {
      Shape4 s(ts[k]->shape());
      ops.p[k] = ts[k]->data();
      for(int j = 0; j < 4; ++j)
        ops.bst[k][j] = s.bst[j];
    }
    ew::GenericGeom g;
    g.out = outS;
    g.total = 1;
    for(int j = 0; j < 4; ++j) {
      g.len[j] = full.d[j] / outS.d[j];
      g.total *= g.len[j];
    }
    g.outLength = outS.elements();
    int blocksX = (g.outLength + 127) / 128;
    // enough slices to fill the machine (148 SMs x 8 blocks), at least 8 reduced elements per slice
    int slices = std::max(1, std::min((kNumSMs * 8 + blocksX - 1) / blocksX, (g.total + 7) / 8));
    slices = std::min(slices, 65535);
    g.chunk = (g.total + slices - 1) / slices;
    slices = (g.total + g.chunk - 1) / g.chunk;
    g.atomic = slices > 1;
    g.assign = (!g.atomic && out->takeLazyZero()) ? 1 : 0;
    dim3 grid(blocksX, slices);
    launchPdl(ew::gAddGeneric<K, Functor>, grid, dim3(128), 0, stream, functor, out->data(), ops, g, scale);
  }
  CUDA_LAUNCH_CHECK();
}

template <class Functor, class... Tensors>
void Add(Functor functor, Tensor out, Tensors... tensors) {
  Add(functor, 1.f, out, tensors...);
}

template <class Functor, class... Tensors>
void Reduce(Functor functor, float scale, Tensor out, Tensors... tensors) {
  out->setLazyZero();  // the accumulating kernel assigns where it can (no fill pass); memset-then-add on the CPU oracle
  Add(functor, scale, out, tensors...);
  out->data();         // (a path that left the mark untouched: materialise)
}

template <class Functor, class... Tensors>
void Reduce(Functor functor, Tensor out, Tensors... tensors) {
  Reduce(functor, 1.f, out, tensors...);
}

}  // namespace marian
