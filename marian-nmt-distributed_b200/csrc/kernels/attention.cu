// Fused multi-head scaled-dot-product attention (forward and backward).
//
// The reference's Transformer::MultiHead / Attention (src/models/transformer.h:153-261) issues,
// per attention block, SplitHeads x3 (reshape + TransposeND), bdot(q, k^T) (cublasSgemmStridedBatched),
// an Element add of the -99999999 mask, Softmax, bdot(weights, v), JoinHeads (TransposeND) - nine
// kernels forward and about twenty backward, with the [B,H,Tq,Tk] scores and their gradient
// round-tripping through HBM several times.  Here one CTA owns one (sentence, head) pair:
//
//   forward : Q,K,V head slices are read straight out of the [B,T,H*dk] projections (the head
//             split is an address computation), S = scale QK^T + mask, row softmax and PV all
//             happen in shared memory; the context is written back in [B,Tq,H*dk] layout (the
//             head join) and the probabilities P are saved for the backward pass.
//   backward: dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(dO o O)), dQ = scale dS K,
//             dK = scale dS^T Q - all five products on shared-memory tiles of one head.
//
// At the config-B shape (B=64, H=8, T=50, dk=64) a head is 3 x 12.8 KB of operands and
// 2 x 160 KFMA per product: far below one tensor-core tile, so the products run as
// register-tiled fp32 FMAs (4x4 micro-tiles, 128-bit shared-memory loads).  That also makes
// the block exact fp32 in every GEMM mode.  Algorithmic HBM bytes: forward reads Q,K,V
// (3 B T d 4) and writes O and P; backward reads Q,K,V,O,dO,P and updates dQ,dK,dV.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "kernels/cuda_helpers.h"
#include "kernels/shadow.h"
#include "kernels/tensor_operators.h"

namespace marian {

namespace {

struct AttnGeom {
  int B, H, Tq, Tk, dk;  // model width d = H * dk
  int maskRows;          // 1 (key mask, broadcast over queries) or Tq (e.g. causal)
  float scale;
};

constexpr int kAttnThreads = 256;

__host__ __device__ inline int pad4(int x) {
  return (x + 3) & ~3;
}

// Copies the [T, dk] head slice of a [B, T, H*dk] tensor into shared memory with row pitch ld.
__device__ __forceinline__ void loadHead(float* dst, int ld, const float* src, int T, int dk, int d) {
  int v4 = dk >> 2;
  for(int e = threadIdx.x; e < T * v4; e += blockDim.x) {
    int r = e / v4, c = (e - r * v4) << 2;
    *reinterpret_cast<float4*>(dst + r * ld + c) = *reinterpret_cast<const float4*>(src + (size_t)r * d + c);
  }
}

// out(i, j) = sum_c A[i*lda + c] * Bm[j*ldb + c]   (i < M, j < N, c < L, L % 4 == 0).
// 4 x 4 outputs per thread; a thread's rows/columns are INTERLEAVED (i = ty + RG q) so that
// neighbouring threads read neighbouring shared-memory rows: with a pitch of 4 (mod 32) words
// the 128-bit loads of a quarter warp fall on distinct banks.
template <class Store>
__device__ __forceinline__ void tileMulNT(const float* A, int lda, int M, const float* Bm, int ldb, int N, int L, Store store) {
  const int RG = (M + 3) >> 2, CG = (N + 3) >> 2;
  for(int tile = threadIdx.x; tile < RG * CG; tile += blockDim.x) {
    const int ty = tile / CG, tx = tile - ty * CG;
    const float* ap[4];
    const float* bp[4];
#pragma unroll
    for(int q = 0; q < 4; ++q) {
      ap[q] = A + min(ty + RG * q, M - 1) * lda;
      bp[q] = Bm + min(tx + CG * q, N - 1) * ldb;
    }
    float acc[4][4];
#pragma unroll
    for(int q = 0; q < 4; ++q)
#pragma unroll
      for(int s = 0; s < 4; ++s)
        acc[q][s] = 0.f;
    for(int c = 0; c < L; c += 4) {
      float4 a[4], b[4];
#pragma unroll
      for(int q = 0; q < 4; ++q) {
        a[q] = *reinterpret_cast<const float4*>(ap[q] + c);
        b[q] = *reinterpret_cast<const float4*>(bp[q] + c);
      }
#pragma unroll
      for(int q = 0; q < 4; ++q)
#pragma unroll
        for(int s = 0; s < 4; ++s) {
          acc[q][s] = fmaf(a[q].x, b[s].x, acc[q][s]);
          acc[q][s] = fmaf(a[q].y, b[s].y, acc[q][s]);
          acc[q][s] = fmaf(a[q].z, b[s].z, acc[q][s]);
          acc[q][s] = fmaf(a[q].w, b[s].w, acc[q][s]);
        }
    }
#pragma unroll
    for(int q = 0; q < 4; ++q) {
      int i = ty + RG * q;
      if(i < M) {
#pragma unroll
        for(int s = 0; s < 4; ++s) {
          int j = tx + CG * s;
          if(j < N)
            store(i, j, acc[q][s]);
        }
      }
    }
  }
}

// out(r, 4t .. 4t+3) = sum_j Pm[r*sr + j*sj] * Vm[j*ldv + 4t ..]   (r < R, t < C/4, j < J).
// sr = ld, sj = 1 multiplies by Pm; sr = 1, sj = ld multiplies by Pm^T.
template <class Store>
__device__ __forceinline__ void tileMulVec(const float* Pm, int sr, int sj, int R, const float* Vm, int ldv, int J, int C, Store store) {
  const int RG = (R + 3) >> 2, C4 = C >> 2;
  for(int tile = threadIdx.x; tile < RG * C4; tile += blockDim.x) {
    const int ty = tile / C4, tx = tile - ty * C4;
    const float* pp[4];
#pragma unroll
    for(int q = 0; q < 4; ++q)
      pp[q] = Pm + min(ty + RG * q, R - 1) * sr;
    float4 acc[4];
#pragma unroll
    for(int q = 0; q < 4; ++q)
      acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* vp = Vm + 4 * tx;
#pragma unroll 2
    for(int j = 0; j < J; ++j) {
      float4 v = *reinterpret_cast<const float4*>(vp + j * ldv);
#pragma unroll
      for(int q = 0; q < 4; ++q) {
        float p = pp[q][j * sj];
        acc[q].x = fmaf(p, v.x, acc[q].x);
        acc[q].y = fmaf(p, v.y, acc[q].y);
        acc[q].z = fmaf(p, v.z, acc[q].z);
        acc[q].w = fmaf(p, v.w, acc[q].w);
      }
    }
#pragma unroll
    for(int q = 0; q < 4; ++q) {
      int r = ty + RG * q;
      if(r < R)
        store(r, 4 * tx, acc[q]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads) gAttentionForward(float* __restrict__ out,
                                                                  float* __restrict__ probs,
                                                                  const float* __restrict__ q,
                                                                  const float* __restrict__ k,
                                                                  const float* __restrict__ v,
                                                                  const float* __restrict__ mask,
                                                                  AttnGeom g) {
  extern __shared__ __align__(16) float smemF[];
  const int ld = g.dk + 4;        // operand row pitch: 4 (mod 32) words for dk = 64
  const int ldS = pad4(g.Tk) + 1;  // odd pitch: the column walk of P^T products stays conflict free
  float* sQ = smemF;
  float* sK = sQ + g.Tq * ld;
  float* sV = sK + g.Tk * ld;
  float* sS = sV + g.Tk * ld;

  const int b = blockIdx.x / g.H, h = blockIdx.x - b * g.H;
  const int d = g.H * g.dk;
  loadHead(sQ, ld, q + ((size_t)b * g.Tq) * d + h * g.dk, g.Tq, g.dk, d);
  loadHead(sK, ld, k + ((size_t)b * g.Tk) * d + h * g.dk, g.Tk, g.dk, d);
  loadHead(sV, ld, v + ((size_t)b * g.Tk) * d + h * g.dk, g.Tk, g.dk, d);
  __syncthreads();

  // S = scale Q K^T + mask
  const float* mrow = mask ? mask + (size_t)b * g.maskRows * g.Tk : nullptr;
  const int maskPitch = g.maskRows > 1 ? g.Tk : 0;
  tileMulNT(sQ, ld, g.Tq, sK, ld, g.Tk, g.dk, [&](int i, int j, float acc) {
    float s = acc * g.scale;
    if(mrow)
      s += mrow[i * maskPitch + j];
    sS[i * ldS + j] = s;
  });
  __syncthreads();

  // row softmax (one warp per row), P saved for the backward pass
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for(int i = warp; i < g.Tq; i += nwarps) {
    float* row = sS + i * ldS;
    float m = -3.0e38f;
    for(int j = lane; j < g.Tk; j += 32)
      m = fmaxf(m, row[j]);
    m = warpMax(m);
    float sum = 0.f;
    for(int j = lane; j < g.Tk; j += 32) {
      float e = __expf(row[j] - m);
      row[j] = e;
      sum += e;
    }
    sum = warpSum(sum);
    float inv = 1.f / sum;
    float* prow = probs ? probs + (((size_t)b * g.H + h) * g.Tq + i) * g.Tk : nullptr;
    for(int j = lane; j < g.Tk; j += 32) {
      float p = row[j] * inv;
      row[j] = p;
      if(prow)
        prow[j] = p;
    }
  }
  __syncthreads();

  // O = P V, written in [B, Tq, H*dk] layout (head join)
  float* ob = out + ((size_t)b * g.Tq) * d + h * g.dk;
  tileMulVec(sS, ldS, 1, g.Tq, sV, ld, g.Tk, g.dk, [&](int i, int c, float4 acc) { *reinterpret_cast<float4*>(ob + (size_t)i * d + c) = acc; });
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit(float* p, float4 v, bool accumulate) {
  if(accumulate) {
    float4 o = *reinterpret_cast<const float4*>(p);
    v.x += o.x;
    v.y += o.y;
    v.z += o.z;
    v.w += o.w;
  }
  *reinterpret_cast<float4*>(p) = v;
}

__global__ void __launch_bounds__(kAttnThreads) gAttentionBackward(float* __restrict__ dq,
                                                                   float* __restrict__ dk_,
                                                                   float* __restrict__ dv,
                                                                   const float* __restrict__ dout,
                                                                   const float* __restrict__ out,
                                                                   const float* __restrict__ probs,
                                                                   const float* __restrict__ q,
                                                                   const float* __restrict__ k,
                                                                   const float* __restrict__ v,
                                                                   AttnGeom g,
                                                                   int accQ,
                                                                   int accK,
                                                                   int accV) {
  extern __shared__ __align__(16) float smemF[];
  const int ld = g.dk + 4;
  const int ldS = pad4(g.Tk) + 1;
  float* sQ = smemF;
  float* sK = sQ + g.Tq * ld;
  float* sV = sK + g.Tk * ld;
  float* sdO = sV + g.Tk * ld;
  float* sP = sdO + g.Tq * ld;  // P, later overwritten by dS
  float* sD = sP + g.Tq * ldS;  // D_i = sum_c dO_ic O_ic = sum_j dP_ij P_ij

  const int b = blockIdx.x / g.H, h = blockIdx.x - b * g.H;
  const int d = g.H * g.dk;
  const size_t offQ = ((size_t)b * g.Tq) * d + h * g.dk;
  const size_t offK = ((size_t)b * g.Tk) * d + h * g.dk;
  loadHead(sQ, ld, q + offQ, g.Tq, g.dk, d);
  loadHead(sK, ld, k + offK, g.Tk, g.dk, d);
  loadHead(sV, ld, v + offK, g.Tk, g.dk, d);
  loadHead(sdO, ld, dout + offQ, g.Tq, g.dk, d);
  const float* pb = probs + (((size_t)b * g.H + h) * g.Tq) * g.Tk;
  for(int e = threadIdx.x; e < g.Tq * g.Tk; e += blockDim.x) {
    int i = e / g.Tk, j = e - i * g.Tk;
    sP[i * ldS + j] = pb[e];
  }
  __syncthreads();

  // D_i (one warp per row; O read from global memory, dO from shared memory)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for(int i = warp; i < g.Tq; i += nwarps) {
    const float* orow = out + offQ + (size_t)i * d;
    float s = 0.f;
    for(int c = lane; c < g.dk; c += 32)
      s = fmaf(sdO[i * ld + c], orow[c], s);
    s = warpSum(s);
    if(lane == 0)
      sD[i] = s;
  }

  // dV = P^T dO
  tileMulVec(sP, 1, ldS, g.Tk, sdO, ld, g.Tq, g.dk, [&](int j, int c, float4 acc) { emit(dv + offK + (size_t)j * d + c, acc, accV != 0); });
  __syncthreads();

  // dS = P o (dO V^T - D), in place over P (each element is read and rewritten by its owner)
  tileMulNT(sdO, ld, g.Tq, sV, ld, g.Tk, g.dk, [&](int i, int j, float acc) {
    float p = sP[i * ldS + j];
    sP[i * ldS + j] = p * (acc - sD[i]) * g.scale;
  });
  __syncthreads();

  // dQ = dS K,  dK = dS^T Q   (scale already folded into dS)
  tileMulVec(sP, ldS, 1, g.Tq, sK, ld, g.Tk, g.dk, [&](int i, int c, float4 acc) { emit(dq + offQ + (size_t)i * d + c, acc, accQ != 0); });
  tileMulVec(sP, 1, ldS, g.Tk, sQ, ld, g.Tq, g.dk, [&](int j, int c, float4 acc) { emit(dk_ + offK + (size_t)j * d + c, acc, accK != 0); });
}


// ------------------------------------------------------------------------------------------
// tensor-core variant: the five products of a head on mma.sync m16n8k8 tf32
// ------------------------------------------------------------------------------------------
// A head's products are 64x64x64-class: too small for a tcgen05 tile pipeline, but a good fit
// for warp-level MMAs on shared-memory operands.  Every matrix is zero-padded to multiples of 16
// in shared memory; row pitches are chosen so that the fragment loads of the m16n8k8 layout are
// bank-conflict free:  A(m,k) / B(k,n) read "along k"  -> pitch = 4 (mod 32) words,
//                      B(k,n) read from a [k][n] matrix -> pitch = 8 (mod 32) words.
// Transposed A operands (P^T, dS^T) are kept as explicit transposed copies instead.
// X3 = true: 3xTF32 error compensation (a = hi + lo, three MMAs) - fp32-grade accuracy for the
// exact GEMM modes; X3 = false: plain tf32, the same operand precision as the tf32 GEMMs.
__device__ __forceinline__ uint32_t toTf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mmaTf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__host__ __device__ inline int pad16(int x) {
  return (x + 15) & ~15;
}
// smallest pitch >= n that is congruent to r modulo 32
__host__ __device__ inline int pitchMod32(int n, int r) {
  int p = (n / 32) * 32 + r;
  return p >= n ? p : p + 32;
}

// D(M x N) = A(M x K) B(K x N); A(m,k) = As[m*lda + k]; B(k,n) = BT ? Bs[n*ldb + k] : Bs[k*ldb + n].
// M % 16 == 0, N % 8 == 0, K % 8 == 0 (zero padded).  Work items = (m-tile, chunk of <= 4 n-tiles),
// dealt round-robin to the warps; the A fragment of a k-step is reused by the chunk's n-tiles.
// epi(row, col, v0, v1) receives the results for (row, col) and (row, col + 1).
template <bool BT, bool X3, class Epi>
__device__ __forceinline__ void warpMmaProduct(const float* As, int lda, const float* Bs, int ldb, int M, int N, int K, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int MT = M >> 4, NT = N >> 3, NC = (NT + 3) >> 2;
  for(int item = warp; item < MT * NC; item += nwarps) {
    const int mt = item % MT, nc = item / MT;
    const int m0 = mt << 4, nt0 = nc << 2;
    const int ntiles = min(4, NT - nt0);
    float acc[4][4];
#pragma unroll
    for(int i = 0; i < 4; ++i)
#pragma unroll
      for(int j = 0; j < 4; ++j)
        acc[i][j] = 0.f;
    const float* ap = As + (m0 + g) * lda + t;
    // B fragment base of n-tile i: &B(k = t, n = n0 + g)
    const float* bp[4];
#pragma unroll
    for(int i = 0; i < 4; ++i) {
      const int n0 = (nt0 + min(i, ntiles - 1)) << 3;
      bp[i] = BT ? Bs + (n0 + g) * ldb + t : Bs + t * ldb + n0 + g;
    }
    const int bk = BT ? 1 : ldb;  // address step of one k
#pragma unroll 2
    for(int k0 = 0; k0 < K; k0 += 8) {
      // all fragment loads of the k-step first (10-12 independent LDS in flight), then the MMAs
      float af[4] = {ap[k0], ap[k0 + 8 * lda], ap[k0 + 4], ap[k0 + 8 * lda + 4]};
      float bf[4][2];
#pragma unroll
      for(int i = 0; i < 4; ++i) {
        bf[i][0] = bp[i][k0 * bk];
        bf[i][1] = bp[i][(k0 + 4) * bk];
      }
      uint32_t ahi[4], alo[4];
#pragma unroll
      for(int i = 0; i < 4; ++i) {
        ahi[i] = toTf32(af[i]);
        if(X3)
          alo[i] = toTf32(af[i] - __uint_as_float(ahi[i]));
      }
#pragma unroll
      for(int i = 0; i < 4; ++i) {
        if(i < ntiles) {
          uint32_t bhi[2] = {toTf32(bf[i][0]), toTf32(bf[i][1])};
          if(X3) {
            uint32_t blo[2] = {toTf32(bf[i][0] - __uint_as_float(bhi[0])), toTf32(bf[i][1] - __uint_as_float(bhi[1]))};
            mmaTf32(acc[i], alo, bhi);  // small terms first
            mmaTf32(acc[i], ahi, blo);
          }
          mmaTf32(acc[i], ahi, bhi);
        }
      }
    }
#pragma unroll
    for(int i = 0; i < 4; ++i) {
      if(i < ntiles) {
        const int col = ((nt0 + i) << 3) + 2 * t;
        epi(m0 + g, col, acc[i][0], acc[i][1]);
        epi(m0 + g + 8, col, acc[i][2], acc[i][3]);
      }
    }
  }
}

// e / n and e % n with a shift when n is a power of two (dk / 4 = 16 and the padded sequence
// lengths 64 / 128 are): integer division costs ~20 instructions per element otherwise
struct FastDiv {
  int n, sh;
  bool p2;
  __device__ __forceinline__ explicit FastDiv(int n_) : n(n_), sh(31 - __builtin_clz((unsigned)n_)), p2((1 << (31 - __builtin_clz((unsigned)n_))) == n_) {}
  __device__ __forceinline__ int div(int e) const { return p2 ? (e >> sh) : (e / n); }
  __device__ __forceinline__ int mod(int e, int q) const { return p2 ? (e & (n - 1)) : (e - q * n); }
};

// Asynchronous global -> shared copies (LDGSTS): every thread fires all of its 16-byte pieces
// and waits ONCE, so a CTA pays one memory round trip for Q, K, V (and dO, P) instead of one per
// loop iteration of a load/store chain.
__device__ __forceinline__ void cpAsync16(float* smem, const float* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsync4(float* smem, const float* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cpAsyncWaitGroup() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cpAsyncWaitAll() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// [T, dk] head slice -> shared memory [Tpad][ld] (asynchronous), rows >= T zero filled
__device__ __forceinline__ void loadHeadPadded(float* dst, int ld, const float* src, int T, int Tpad, int dk, int d) {
  const int v4 = dk >> 2;
  const FastDiv fd(v4);
  for(int e = threadIdx.x; e < Tpad * v4; e += blockDim.x) {
    int r = fd.div(e), c = fd.mod(e, r) << 2;
    if(r < T)
      cpAsync16(dst + r * ld + c, src + (size_t)r * d + c);
    else
      *reinterpret_cast<float4*>(dst + r * ld + c) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// shared memory [T][ld] -> [T, dk] head slice of a [B, T, H*dk] tensor, 128-bit coalesced rows
__device__ __forceinline__ void storeHead(float* dst, const float* src, int ld, int T, int dk, int d, bool accumulate) {
  const int v4 = dk >> 2;
  const FastDiv fd(v4);
  for(int e = threadIdx.x; e < T * v4; e += blockDim.x) {
    int r = fd.div(e), c = fd.mod(e, r) << 2;
    float4 v = *reinterpret_cast<const float4*>(src + r * ld + c);
    float* gp = dst + (size_t)r * d + c;
    if(accumulate) {
      float4 o = *reinterpret_cast<const float4*>(gp);
      v.x += o.x;
      v.y += o.y;
      v.z += o.z;
      v.w += o.w;
    }
    *reinterpret_cast<float4*>(gp) = v;
  }
}

// Forward shared-memory plan: three tiles only, so that FOUR CTAs fit on an SM and the 512 heads
// of a config-B block run as a single wave:
//   tile A: Q, later reused for V (loaded asynchronously WHILE the softmax runs)
//   tile B: K, later the staging tile of the output O
//   tile S: scores / probabilities
struct MmaLayoutFwd {
  int TqP, TkP, rowsA, rowsB, ldQ, ldK, ldV, ldS;
  __host__ __device__ MmaLayoutFwd(const AttnGeom& g) {
    TqP = pad16(g.Tq);
    TkP = pad16(g.Tk);
    rowsA = TqP > TkP ? TqP : TkP;
    rowsB = rowsA;
    ldQ = ldK = pitchMod32(g.dk, 4);
    ldV = pitchMod32(g.dk, 8);
    ldS = pitchMod32(TkP, 4);
  }
  __host__ __device__ size_t floats() const { return (size_t)rowsA * ldV + (size_t)rowsB * ldK + (size_t)TqP * ldS; }
};

template <bool X3>
__global__ void __launch_bounds__(kAttnThreads) gAttentionForwardMma(float* __restrict__ out,
                                                                     float* __restrict__ probs,
                                                                     const float* __restrict__ q,
                                                                     const float* __restrict__ k,
                                                                     const float* __restrict__ v,
                                                                     const float* __restrict__ mask,
                                                                     AttnGeom g) {
  extern __shared__ __align__(16) float smemF[];
  pdlEnter();
  const MmaLayoutFwd L(g);
  float* sQ = smemF;                 // tile A (pitch ldQ while it holds Q, ldV while it holds V)
  float* sV = smemF;
  float* sK = sQ + L.rowsA * L.ldV;  // tile B
  float* sO = sK;
  float* sS = sK + L.rowsB * L.ldK;

  const int b = blockIdx.x / g.H, h = blockIdx.x - b * g.H;
  const int d = g.H * g.dk;
  loadHeadPadded(sQ, L.ldQ, q + ((size_t)b * g.Tq) * d + h * g.dk, g.Tq, L.TqP, g.dk, d);
  loadHeadPadded(sK, L.ldK, k + ((size_t)b * g.Tk) * d + h * g.dk, g.Tk, L.TkP, g.dk, d);
  cpAsyncWaitAll();
  __syncthreads();

  // S = scale Q K^T + mask
  const float* mrow = mask ? mask + (size_t)b * g.maskRows * g.Tk : nullptr;
  const int maskPitch = g.maskRows > 1 ? g.Tk : 0;
  warpMmaProduct<true, X3>(sQ, L.ldQ, sK, L.ldK, L.TqP, L.TkP, g.dk, [&](int i, int j, float v0, float v1) {
    float s0 = v0 * g.scale, s1 = v1 * g.scale;
    if(mrow && i < g.Tq) {
      if(j < g.Tk)
        s0 += mrow[i * maskPitch + j];
      if(j + 1 < g.Tk)
        s1 += mrow[i * maskPitch + j + 1];
    }
    *reinterpret_cast<float2*>(sS + i * L.ldS + j) = make_float2(s0, s1);
  });
  __syncthreads();

  // Q is dead: V streams into its tile while the softmax runs
  loadHeadPadded(sV, L.ldV, v + ((size_t)b * g.Tk) * d + h * g.dk, g.Tk, L.TkP, g.dk, d);

  // row softmax over the real columns; padding columns / rows become exact zeros
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for(int i = warp; i < L.TqP; i += nwarps) {
    float* row = sS + i * L.ldS;
    if(i >= g.Tq) {
      for(int j = lane; j < L.TkP; j += 32)
        row[j] = 0.f;
      continue;
    }
    float m = -3.0e38f;
    for(int j = lane; j < g.Tk; j += 32)
      m = fmaxf(m, row[j]);
    m = warpMax(m);
    float sum = 0.f;
    for(int j = lane; j < g.Tk; j += 32) {
      float e = __expf(row[j] - m);
      row[j] = e;
      sum += e;
    }
    sum = warpSum(sum);
    float inv = 1.f / sum;
    float* prow = probs ? probs + (((size_t)b * g.H + h) * g.Tq + i) * g.Tk : nullptr;
    for(int j = lane; j < L.TkP; j += 32) {
      float p = j < g.Tk ? row[j] * inv : 0.f;
      row[j] = p;
      if(prow && j < g.Tk)
        prow[j] = p;
    }
  }
  cpAsyncWaitAll();
  __syncthreads();

  // O = P V, staged in the (now idle) K tile and written as coalesced rows of [B, Tq, H*dk]
  warpMmaProduct<false, X3>(sS, L.ldS, sV, L.ldV, L.TqP, g.dk, L.TkP, [&](int i, int c, float v0, float v1) {
    *reinterpret_cast<float2*>(sO + i * L.ldK + c) = make_float2(v0, v1);
  });
  __syncthreads();
  storeHead(out + ((size_t)b * g.Tq) * d + h * g.dk, sO, L.ldK, g.Tq, g.dk, d, false);
}

// ------------------------------------------------------------------------------------------
// forward, warp-private formulation: ONE block barrier
// ------------------------------------------------------------------------------------------
// A CTA of 4 warps owns 64 query rows of one (sentence, head); warp w owns rows 16w..16w+15:
//   S strip (16 x Tk) = Q_w K^T stays in the warp's accumulator registers, the row softmax runs on
//   those registers (a row lives in the 4 lanes of a quad: two shuffles per reduction), P goes to
//   global memory (for the backward pass) and into the warp's OWN 16 rows of the Q tile (Q is dead
//   after the strip: its A fragments were consumed), O strip = P_w V, written out by the warp.
// K and V are the only data shared by the warps -> one __syncthreads after the cp.async loads,
// afterwards only __syncwarp.  Three shared-memory tiles (Q/P, K, V) -> 4 CTAs per SM.
// NTS = number of 8-wide score tiles the registers are sized for (Tk <= 8 NTS), dk <= 64.
template <bool X3>
__device__ __forceinline__ void mmaSplit(float (&d)[4], const float (&af)[4], const uint32_t (&ahi)[4], float b0, float b1) {
  uint32_t bhi[2] = {toTf32(b0), toTf32(b1)};
  if(X3) {
    uint32_t alo[4], blo[2] = {toTf32(b0 - __uint_as_float(bhi[0])), toTf32(b1 - __uint_as_float(bhi[1]))};
#pragma unroll
    for(int i = 0; i < 4; ++i)
      alo[i] = toTf32(af[i] - __uint_as_float(ahi[i]));
    mmaTf32(d, alo, bhi);
    mmaTf32(d, ahi, blo);
  }
  mmaTf32(d, ahi, bhi);
}

template <bool X3, int NTS>
__global__ void __launch_bounds__(128) gAttentionForwardWarp(float* __restrict__ out,
                                                             float* __restrict__ probs,
                                                             const float* __restrict__ q,
                                                             const float* __restrict__ k,
                                                             const float* __restrict__ v,
                                                             const float* __restrict__ mask,
                                                             AttnGeom g) {
  extern __shared__ __align__(16) float smemF[];
  pdlEnter();
  const int TkP = pad16(g.Tk);
  const int ldQ = pitchMod32(g.dk > TkP ? g.dk : TkP, 4);  // Q rows are later overwritten by P rows (TkP wide)
  const int ldK = pitchMod32(g.dk, 4), ldV = pitchMod32(g.dk, 8);
  float* sQ = smemF;            // [64][ldQ]
  float* sK = sQ + 64 * ldQ;    // [TkP][ldK]
  float* sV = sK + TkP * ldK;   // [TkP][ldV]

  const int bh = blockIdx.x, b = bh / g.H, h = bh - b * g.H;
  const int row0 = blockIdx.y * 64;  // first query row of this CTA
  const int d = g.H * g.dk;
  const int rowsHere = min(64, g.Tq - row0);
  loadHeadPadded(sQ, ldQ, q + ((size_t)b * g.Tq + row0) * d + h * g.dk, rowsHere, 64, g.dk, d);
  loadHeadPadded(sK, ldK, k + ((size_t)b * g.Tk) * d + h * g.dk, g.Tk, TkP, g.dk, d);
  loadHeadPadded(sV, ldV, v + ((size_t)b * g.Tk) * d + h * g.dk, g.Tk, TkP, g.dk, d);
  cpAsyncWaitAll();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  if(m0 >= rowsHere)
    return;  // whole strip is padding (no further block barriers below)
  const int nts = TkP >> 3;

  // ---- S strip = Q_w K^T ----
  float acc[NTS][4];
#pragma unroll
  for(int i = 0; i < NTS; ++i)
    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  {
    const float* ap = sQ + (m0 + gq) * ldQ + t;
    for(int k0 = 0; k0 < g.dk; k0 += 8) {
      float af[4] = {ap[k0], ap[k0 + 8 * ldQ], ap[k0 + 4], ap[k0 + 8 * ldQ + 4]};
      uint32_t ahi[4] = {toTf32(af[0]), toTf32(af[1]), toTf32(af[2]), toTf32(af[3])};
#pragma unroll
      for(int i = 0; i < NTS; ++i)
        if(i < nts) {
          const float* bp = sK + (i * 8 + gq) * ldK + k0 + t;
          mmaSplit<X3>(acc[i], af, ahi, bp[0], bp[4]);
        }
    }
  }
  __syncwarp();  // every lane of the warp is done reading its Q rows: they become the P strip

  // ---- scale, mask, row softmax on the accumulators (rows gq and gq + 8 of the strip) ----
  const int iLo = row0 + m0 + gq, iHi = iLo + 8;
  const float* mLo = nullptr;
  const float* mHi = nullptr;
  if(mask) {
    const float* mb = mask + (size_t)b * g.maskRows * g.Tk;
    mLo = mb + (g.maskRows > 1 ? (size_t)min(iLo, g.Tq - 1) * g.Tk : 0);
    mHi = mb + (g.maskRows > 1 ? (size_t)min(iHi, g.Tq - 1) * g.Tk : 0);
  }
  float mxLo = -3.0e38f, mxHi = -3.0e38f;
#pragma unroll
  for(int i = 0; i < NTS; ++i)
    if(i < nts) {
#pragma unroll
      for(int e = 0; e < 2; ++e) {
        int j = i * 8 + 2 * t + e;
        bool real = j < g.Tk;
        float lo = real ? acc[i][e] * g.scale + (mLo ? mLo[j] : 0.f) : -3.0e38f;
        float hi = real ? acc[i][2 + e] * g.scale + (mHi ? mHi[j] : 0.f) : -3.0e38f;
        acc[i][e] = lo;
        acc[i][2 + e] = hi;
        mxLo = fmaxf(mxLo, lo);
        mxHi = fmaxf(mxHi, hi);
      }
    }
  mxLo = fmaxf(mxLo, __shfl_xor_sync(0xffffffffu, mxLo, 1));
  mxLo = fmaxf(mxLo, __shfl_xor_sync(0xffffffffu, mxLo, 2));
  mxHi = fmaxf(mxHi, __shfl_xor_sync(0xffffffffu, mxHi, 1));
  mxHi = fmaxf(mxHi, __shfl_xor_sync(0xffffffffu, mxHi, 2));
  float sumLo = 0.f, sumHi = 0.f;
#pragma unroll
  for(int i = 0; i < NTS; ++i)
    if(i < nts) {
#pragma unroll
      for(int e = 0; e < 2; ++e) {
        int j = i * 8 + 2 * t + e;
        float lo = j < g.Tk ? __expf(acc[i][e] - mxLo) : 0.f;
        float hi = j < g.Tk ? __expf(acc[i][2 + e] - mxHi) : 0.f;
        acc[i][e] = lo;
        acc[i][2 + e] = hi;
        sumLo += lo;
        sumHi += hi;
      }
    }
  sumLo += __shfl_xor_sync(0xffffffffu, sumLo, 1);
  sumLo += __shfl_xor_sync(0xffffffffu, sumLo, 2);
  sumHi += __shfl_xor_sync(0xffffffffu, sumHi, 1);
  sumHi += __shfl_xor_sync(0xffffffffu, sumHi, 2);
  const float invLo = 1.f / sumLo, invHi = 1.f / sumHi;

  // ---- P strip -> global (backward pass) and -> the warp's rows of the Q tile (A operand of P V) ----
  float* sP = sQ + m0 * ldQ;
  float* pLo = (probs && iLo < g.Tq) ? probs + (((size_t)b * g.H + h) * g.Tq + iLo) * g.Tk : nullptr;
  float* pHi = (probs && iHi < g.Tq) ? probs + (((size_t)b * g.H + h) * g.Tq + iHi) * g.Tk : nullptr;
#pragma unroll
  for(int i = 0; i < NTS; ++i)
    if(i < nts) {
      int j = i * 8 + 2 * t;
      float l0 = acc[i][0] * invLo, l1 = acc[i][1] * invLo, h0 = acc[i][2] * invHi, h1 = acc[i][3] * invHi;
      *reinterpret_cast<float2*>(sP + gq * ldQ + j) = make_float2(l0, l1);
      *reinterpret_cast<float2*>(sP + (gq + 8) * ldQ + j) = make_float2(h0, h1);
      if(pLo) {
        if(j < g.Tk)
          pLo[j] = l0;
        if(j + 1 < g.Tk)
          pLo[j + 1] = l1;
      }
      if(pHi) {
        if(j < g.Tk)
          pHi[j] = h0;
        if(j + 1 < g.Tk)
          pHi[j + 1] = h1;
      }
    }
  __syncwarp();

  // ---- O strip = P_w V ----
  const int nto = g.dk >> 3;
  float o[8][4];
#pragma unroll
  for(int i = 0; i < 8; ++i)
    o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  {
    const float* ap = sP + gq * ldQ + t;
    for(int k0 = 0; k0 < TkP; k0 += 8) {
      float af[4] = {ap[k0], ap[k0 + 8 * ldQ], ap[k0 + 4], ap[k0 + 8 * ldQ + 4]};
      uint32_t ahi[4] = {toTf32(af[0]), toTf32(af[1]), toTf32(af[2]), toTf32(af[3])};
#pragma unroll
      for(int i = 0; i < 8; ++i)
        if(i < nto) {
          const float* bp = sV + (k0 + t) * ldV + i * 8 + gq;
          mmaSplit<X3>(o[i], af, ahi, bp[0], bp[4 * ldV]);
        }
    }
  }
  float* oLo = iLo < g.Tq ? out + ((size_t)b * g.Tq + iLo) * d + h * g.dk : nullptr;
  float* oHi = iHi < g.Tq ? out + ((size_t)b * g.Tq + iHi) * d + h * g.dk : nullptr;
#pragma unroll
  for(int i = 0; i < 8; ++i)
    if(i < nto) {
      int c = i * 8 + 2 * t;
      if(oLo)
        *reinterpret_cast<float2*>(oLo + c) = make_float2(o[i][0], o[i][1]);
      if(oHi)
        *reinterpret_cast<float2*>(oHi + c) = make_float2(o[i][2], o[i][3]);
    }
}


// Backward shared-memory plan: four tiles (+ D), three CTAs per SM at the config-B shape:
//   tile 1: dO (phases dV, dP), then Q (phase dK)          tile 2: V (phase dP), then K (phase dQ),
//   tile P: P -> dS                                          tile PT: P^T -> dS^T
// K and Q arrive (cp.async) once dO and V are dead; dQ / dK / dV go straight to global memory.
struct MmaLayoutBwd {
  int TqP, TkP, rows1, rows2, ld1, ld2, ldQ, ldK, ldV, ldO, ldS, ldT;
  __host__ __device__ MmaLayoutBwd(const AttnGeom& g) {
    TqP = pad16(g.Tq);
    TkP = pad16(g.Tk);
    ldQ = ldK = pitchMod32(g.dk, 8);  // B operands read from [k][n] matrices
    ldV = pitchMod32(g.dk, 4);        // B operand read along k
    ldO = pitchMod32(g.dk, 12);       // dO is A along k (conflict free) and B from [k][n] (2-way)
    ldS = pitchMod32(TkP, 4);         // P / dS   [TqP][TkP]
    ldT = pitchMod32(TqP, 4);         // P^T / dS^T [TkP][TqP]
    rows1 = TqP;                      // dO [TqP], Q [TqP]
    ld1 = ldO > ldQ ? ldO : ldQ;
    rows2 = TkP;                      // V [TkP], K [TkP]
    ld2 = ldV > ldK ? ldV : ldK;
  }
  __host__ __device__ size_t floats() const { return (size_t)rows1 * ld1 + (size_t)rows2 * ld2 + (size_t)TqP * ldS + (size_t)TkP * ldT + TqP; }
};

__device__ __forceinline__ void emit2(float* p, float v0, float v1, bool accumulate) {
  float2 v = make_float2(v0, v1);
  if(accumulate) {
    float2 o = *reinterpret_cast<const float2*>(p);
    v.x += o.x;
    v.y += o.y;
  }
  *reinterpret_cast<float2*>(p) = v;
}

template <bool X3>
__global__ void __launch_bounds__(kAttnThreads) gAttentionBackwardMma(float* __restrict__ dq,
                                                                      float* __restrict__ dk_,
                                                                      float* __restrict__ dv,
                                                                      const float* __restrict__ dout,
                                                                      const float* __restrict__ out,
                                                                      const float* __restrict__ probs,
                                                                      const float* __restrict__ q,
                                                                      const float* __restrict__ k,
                                                                      const float* __restrict__ v,
                                                                      AttnGeom g,
                                                                      int accQ,
                                                                      int accK,
                                                                      int accV) {
  extern __shared__ __align__(16) float smemF[];
  pdlEnter();
  const MmaLayoutBwd L(g);
  float* t1 = smemF;                    // dO, later Q
  float* t2 = t1 + L.rows1 * L.ld1;     // V, later K
  float* sP = t2 + L.rows2 * L.ld2;     // P, later dS
  float* sPT = sP + L.TqP * L.ldS;      // P^T, later dS^T
  float* sD = sPT + L.TkP * L.ldT;
  float* sdO = t1;
  float* sV = t2;

  const int b = blockIdx.x / g.H, h = blockIdx.x - b * g.H;
  const int d = g.H * g.dk;
  const size_t offQ = ((size_t)b * g.Tq) * d + h * g.dk;
  const size_t offK = ((size_t)b * g.Tk) * d + h * g.dk;
  loadHeadPadded(sdO, L.ldO, dout + offQ, g.Tq, L.TqP, g.dk, d);
  loadHeadPadded(sV, L.ldV, v + offK, g.Tk, L.TkP, g.dk, d);
  const float* pb = probs + (((size_t)b * g.H + h) * g.Tq) * g.Tk;
  const FastDiv fdk(L.TkP), fdq(L.TqP);
  for(int e = threadIdx.x; e < L.TqP * L.TkP; e += blockDim.x) {
    int i = fdk.div(e), j = fdk.mod(e, i);
    if(i < g.Tq && j < g.Tk)
      cpAsync4(sP + i * L.ldS + j, pb + i * g.Tk + j);
    else
      sP[i * L.ldS + j] = 0.f;
  }
  // the forward output O (only needed for D_i) rides along into the tile that will hold P^T
  float* sO = sPT;
  const int ldOut = g.dk;  // dense rows: dk <= ldT * TkP / TqP is checked by the launcher
  {
    const int v4 = g.dk >> 2;
    const FastDiv fd(v4);
    for(int e = threadIdx.x; e < g.Tq * v4; e += blockDim.x) {
      int r = fd.div(e), c = fd.mod(e, r) << 2;
      cpAsync16(sO + r * ldOut + c, out + offQ + (size_t)r * d + c);
    }
  }
  cpAsyncWaitAll();
  __syncthreads();
  // D_i = sum_c dO_ic O_ic
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for(int i = warp; i < L.TqP; i += nwarps) {
    float s = 0.f;
    if(i < g.Tq)
      for(int c = lane; c < g.dk; c += 32)
        s = fmaf(sdO[i * L.ldO + c], sO[i * ldOut + c], s);
    s = warpSum(s);
    if(lane == 0)
      sD[i] = s;
  }
  __syncthreads();
  // explicit transposed copy (over O): P^T is the A operand of dV = P^T dO
  for(int e = threadIdx.x; e < L.TqP * L.TkP; e += blockDim.x) {
    int j = fdq.div(e), i = fdq.mod(e, j);
    sPT[j * L.ldT + i] = sP[i * L.ldS + j];
  }
  __syncthreads();

  // dV = P^T dO
  warpMmaProduct<false, X3>(sPT, L.ldT, sdO, L.ldO, L.TkP, g.dk, L.TqP, [&](int j, int c, float v0, float v1) {
    if(j < g.Tk)
      emit2(dv + offK + (size_t)j * d + c, v0, v1, accV != 0);
  });
  __syncthreads();

  // dS = scale * P o (dO V^T - D), written over P and P^T (each element by its owner thread)
  warpMmaProduct<true, X3>(sdO, L.ldO, sV, L.ldV, L.TqP, L.TkP, g.dk, [&](int i, int j, float v0, float v1) {
    float2 p = *reinterpret_cast<const float2*>(sP + i * L.ldS + j);
    float di = sD[i];
    float s0 = p.x * (v0 - di) * g.scale, s1 = p.y * (v1 - di) * g.scale;
    *reinterpret_cast<float2*>(sP + i * L.ldS + j) = make_float2(s0, s1);
    sPT[j * L.ldT + i] = s0;
    sPT[(j + 1) * L.ldT + i] = s1;
  });
  __syncthreads();

  // dO and V are dead: K and Q stream into their tiles
  float* sK = t2;
  float* sQ = t1;
  loadHeadPadded(sK, L.ldK, k + offK, g.Tk, L.TkP, g.dk, d);
  loadHeadPadded(sQ, L.ldQ, q + offQ, g.Tq, L.TqP, g.dk, d);
  cpAsyncWaitAll();
  __syncthreads();

  // dQ = dS K,  dK = dS^T Q  (written straight to global memory: no idle tile left to stage in)
  warpMmaProduct<false, X3>(sP, L.ldS, sK, L.ldK, L.TqP, g.dk, L.TkP, [&](int i, int c, float v0, float v1) {
    if(i < g.Tq)
      emit2(dq + offQ + (size_t)i * d + c, v0, v1, accQ != 0);
  });
  warpMmaProduct<false, X3>(sPT, L.ldT, sQ, L.ldQ, L.TkP, g.dk, L.TqP, [&](int j, int c, float v0, float v1) {
    if(j < g.Tk)
      emit2(dk_ + offK + (size_t)j * d + c, v0, v1, accK != 0);
  });
}

// Stores a 16 x dk accumulator strip (rows lo / hi of a lane) to [.., d]-pitched global memory.
// Accumulating: all old values are loaded before the first store (one memory round trip, not 2 nto).
// slo / shi (optional): the same rows of the tensor's bf16 shadow (BF16S GEMM mode), written with the final values.
__device__ __forceinline__ void storeStrip(float* lo, float* hi, float (&o)[8][4], int nto, int t, bool accumulate, __nv_bfloat16* slo = nullptr, __nv_bfloat16* shi = nullptr) {
  if(accumulate) {
#pragma unroll
    for(int n = 0; n < 8; ++n)
      if(n < nto) {
        if(lo) {
          float2 x = *reinterpret_cast<const float2*>(lo + n * 8 + 2 * t);
          o[n][0] += x.x;
          o[n][1] += x.y;
        }
        if(hi) {
          float2 x = *reinterpret_cast<const float2*>(hi + n * 8 + 2 * t);
          o[n][2] += x.x;
          o[n][3] += x.y;
        }
      }
  }
#pragma unroll
  for(int n = 0; n < 8; ++n)
    if(n < nto) {
      if(lo)
        *reinterpret_cast<float2*>(lo + n * 8 + 2 * t) = make_float2(o[n][0], o[n][1]);
      if(hi)
        *reinterpret_cast<float2*>(hi + n * 8 + 2 * t) = make_float2(o[n][2], o[n][3]);
      if(slo)
        *reinterpret_cast<__nv_bfloat162*>(slo + n * 8 + 2 * t) = __floats2bfloat162_rn(o[n][0], o[n][1]);
      if(shi)
        *reinterpret_cast<__nv_bfloat162*>(shi + n * 8 + 2 * t) = __floats2bfloat162_rn(o[n][2], o[n][3]);
    }
}

// ------------------------------------------------------------------------------------------
// backward, warp-private formulation (Tq, Tk <= 64, dk == 64): four tiles, three block barriers
// ------------------------------------------------------------------------------------------
// Phase 1, warp w owns QUERY rows 16w..16w+15, everything in registers:
//   dP strip = dO_w V^T;  P strip straight from global memory in the accumulator layout;
//   D_i = sum_j P_ij dP_ij (quad shuffles);  dS = scale P o (dP - D);
//   dQ strip = dS_w K with the A operand taken FROM THE ACCUMULATOR REGISTERS: the contraction
//   index may be permuted freely, so k-slot t of step i stands for key 8i+2t and slot t+4 for key
//   8i+2t+1 - exactly the two columns a lane holds - and the B fragment reads K rows 8i+2t, 8i+2t+1.
// Phase 2, warp w owns KEY rows 16w..16w+15: dV = P^T dO and dK = dS^T Q read P / dS (stored by the
//   phase-1 owners over the dead V / K tiles) as transposed A operands with the same slot permutation.
// Tiles are [R][64] floats WITHOUT padding, column index XOR-swizzled with 4 (row mod 8): every
// fragment load above is bank-conflict free, 16-byte chunks stay contiguous (cp.async), and at
// R = 56 rows (T <= 56) the four tiles take exactly 56 KB -> FOUR CTAs per SM, so the 512 heads of
// a config-B block run as one wave (with padded 68-float pitches: three CTAs, 444 slots, a tail wave).
// The forward output O is not needed (D_i comes from P and dP).
__device__ __forceinline__ int swz(int row) {
  return (row & 7) << 2;
}

// [T, 64] head slice -> swizzled [R][64] tile (asynchronous), rows >= T zero filled
__device__ __forceinline__ void loadHeadSwizzled(float* dst, const float* src, int T, int R, int d) {
  for(int e = threadIdx.x; e < R * 16; e += blockDim.x) {
    int r = e >> 4, c = (e & 15) << 2;
    float* dp = dst + r * 64 + (c ^ swz(r));
    if(r < T)
      cpAsync16(dp, src + (size_t)r * d + c);
    else
      *reinterpret_cast<float4*>(dp) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <bool X3>
__global__ void __launch_bounds__(128) gAttentionBackwardWarp(float* __restrict__ dq,
                                                              float* __restrict__ dk_,
                                                              float* __restrict__ dv,
                                                              const float* __restrict__ dout,
                                                              const float* __restrict__ probs,
                                                              const float* __restrict__ q,
                                                              const float* __restrict__ k,
                                                              const float* __restrict__ v,
                                                              AttnGeom g,
                                                              int R,
                                                              int accQ,
                                                              int accK,
                                                              int accV,
                                                              __nv_bfloat16* __restrict__ dqS,
                                                              __nv_bfloat16* __restrict__ dkS,
                                                              __nv_bfloat16* __restrict__ dvS) {
  extern __shared__ __align__(16) float smemF[];
  pdlEnter();
  const int TILE = R * 64;
  float* sdO = smemF;
  float* sV = sdO + TILE;
  float* sK = sV + TILE;
  float* sQ = sK + TILE;
  float* sP = sV;   // P over V (after barrier A)
  float* sdS = sK;  // dS over K (after barrier A2)

  const int b = blockIdx.x / g.H, h = blockIdx.x - b * g.H;
  const int d = g.H * 64;
  const size_t offQ = ((size_t)b * g.Tq) * d + h * 64;
  const size_t offK = ((size_t)b * g.Tk) * d + h * 64;
  loadHeadSwizzled(sdO, dout + offQ, g.Tq, R, d);
  loadHeadSwizzled(sV, v + offK, g.Tk, R, d);
  cpAsyncCommit();
  loadHeadSwizzled(sK, k + offK, g.Tk, R, d);
  loadHeadSwizzled(sQ, q + offQ, g.Tq, R, d);
  cpAsyncCommit();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  const int nts = (g.Tk + 7) >> 3;  // 8-wide key tiles that hold real keys
  const int nqs = (g.Tq + 7) >> 3;  // 8-deep query steps that hold real queries
  const int iLo = m0 + gq, iHi = iLo + 8;

  // ---- P strip from global memory, in the accumulator layout (zero outside Tq x Tk) ----
  float p[8][4];
  {
    const float* pb = probs + (((size_t)b * g.H + h) * g.Tq) * g.Tk;
    const bool pair = (g.Tk & 1) == 0;
#pragma unroll
    for(int i = 0; i < 8; ++i) {
      const int j = i * 8 + 2 * t;
      p[i][0] = p[i][1] = p[i][2] = p[i][3] = 0.f;
      if(pair) {
        if(j < g.Tk) {
          if(iLo < g.Tq) {
            float2 x = *reinterpret_cast<const float2*>(pb + (size_t)iLo * g.Tk + j);
            p[i][0] = x.x;
            p[i][1] = x.y;
          }
          if(iHi < g.Tq) {
            float2 x = *reinterpret_cast<const float2*>(pb + (size_t)iHi * g.Tk + j);
            p[i][2] = x.x;
            p[i][3] = x.y;
          }
        }
      } else {
#pragma unroll
        for(int e = 0; e < 2; ++e)
          if(j + e < g.Tk) {
            if(iLo < g.Tq)
              p[i][e] = pb[(size_t)iLo * g.Tk + j + e];
            if(iHi < g.Tq)
              p[i][2 + e] = pb[(size_t)iHi * g.Tk + j + e];
          }
      }
    }
  }
  cpAsyncWaitGroup<1>();
  __syncthreads();  // dO and V have landed

  // rows m0+gq, m0+gq+8 and 8i+gq all have (row mod 8) == gq: one swizzle term for A and NT-B reads
  const int cA = t ^ (gq << 2);
  // ---- dP strip = dO_w V^T ----
  float acc[8][4];
#pragma unroll
  for(int i = 0; i < 8; ++i)
    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  {
    // strips beyond the tile (pure padding, results never stored) re-read rows gq of the tile: the
    // row index keeps (row mod 8) == gq, so the swizzle term is unchanged
    const float* ap = sdO + (iLo < R ? iLo : gq) * 64;
    const int hiOff = ((iHi < R ? iHi : gq) - (iLo < R ? iLo : gq)) * 64;
    const float* bp0 = sV + gq * 64;
#pragma unroll 2
    for(int k0 = 0; k0 < 64; k0 += 8) {
      const int c0 = k0 ^ cA, c1 = c0 ^ 4;
      float af[4] = {ap[c0], ap[c0 + hiOff], ap[c1], ap[c1 + hiOff]};
      uint32_t ahi[4] = {toTf32(af[0]), toTf32(af[1]), toTf32(af[2]), toTf32(af[3])};
#pragma unroll
      for(int i = 0; i < 8; ++i)
        if(i < nts)
          mmaSplit<X3>(acc[i], af, ahi, bp0[i * 512 + c0], bp0[i * 512 + c1]);
    }
  }
  // ---- D_i and dS (over the dP registers) ----
  {
    float dLo = 0.f, dHi = 0.f;
#pragma unroll
    for(int i = 0; i < 8; ++i) {
      dLo = fmaf(p[i][0], acc[i][0], fmaf(p[i][1], acc[i][1], dLo));
      dHi = fmaf(p[i][2], acc[i][2], fmaf(p[i][3], acc[i][3], dHi));
    }
    dLo += __shfl_xor_sync(0xffffffffu, dLo, 1);
    dLo += __shfl_xor_sync(0xffffffffu, dLo, 2);
    dHi += __shfl_xor_sync(0xffffffffu, dHi, 1);
    dHi += __shfl_xor_sync(0xffffffffu, dHi, 2);
#pragma unroll
    for(int i = 0; i < 8; ++i) {
      acc[i][0] = p[i][0] * (acc[i][0] - dLo) * g.scale;
      acc[i][1] = p[i][1] * (acc[i][1] - dLo) * g.scale;
      acc[i][2] = p[i][2] * (acc[i][2] - dHi) * g.scale;
      acc[i][3] = p[i][3] * (acc[i][3] - dHi) * g.scale;
    }
  }
  cpAsyncWaitGroup<0>();
  __syncthreads();  // barrier A: V is dead, K and Q have landed

  // P strip -> V tile (read transposed in phase 2); rows >= R are zero and never read
  {
    float* lo = sP + iLo * 64;
    const int cw = (2 * t) ^ (gq << 2);
#pragma unroll
    for(int i = 0; i < 8; ++i) {
      if(iLo < R)
        *reinterpret_cast<float2*>(lo + ((i * 8) ^ cw)) = make_float2(p[i][0], p[i][1]);
      if(iHi < R)
        *reinterpret_cast<float2*>(lo + 8 * 64 + ((i * 8) ^ cw)) = make_float2(p[i][2], p[i][3]);
    }
  }

  // permuted [k][n] reads: rows 8s+2t (swizzle 8t) and 8s+2t+1 (swizzle 8t+4), column 8n+gq
  const int cE = gq ^ (t << 3), cO = cE ^ 4;
  float o[8][4];
  // ---- dQ strip = dS_w K, A operand from the dS registers ----
#pragma unroll
  for(int n = 0; n < 8; ++n)
    o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for(int i = 0; i < 8; ++i)
    if(i < nts) {
      float af[4] = {acc[i][0], acc[i][2], acc[i][1], acc[i][3]};
      uint32_t ahi[4] = {toTf32(af[0]), toTf32(af[1]), toTf32(af[2]), toTf32(af[3])};
      const float* kp = sK + (i * 8 + 2 * t) * 64;
#pragma unroll
      for(int n = 0; n < 8; ++n)
        mmaSplit<X3>(o[n], af, ahi, kp[(n * 8) ^ cE], kp[64 + ((n * 8) ^ cO)]);
    }
  storeStrip((dq && iLo < g.Tq) ? dq + offQ + (size_t)iLo * d : nullptr, (dq && iHi < g.Tq) ? dq + offQ + (size_t)iHi * d : nullptr, o, 8, t, accQ != 0,
             (dqS && iLo < g.Tq) ? dqS + offQ + (size_t)iLo * d : nullptr, (dqS && iHi < g.Tq) ? dqS + offQ + (size_t)iHi * d : nullptr);
  __syncthreads();  // barrier A2: K is dead, the P tile is complete

  // dS strip -> K tile
  {
    float* lo = sdS + iLo * 64;
    const int cw = (2 * t) ^ (gq << 2);
#pragma unroll
    for(int i = 0; i < 8; ++i) {
      if(iLo < R)
        *reinterpret_cast<float2*>(lo + ((i * 8) ^ cw)) = make_float2(acc[i][0], acc[i][1]);
      if(iHi < R)
        *reinterpret_cast<float2*>(lo + 8 * 64 + ((i * 8) ^ cw)) = make_float2(acc[i][2], acc[i][3]);
    }
  }

  // ---- phase 2: the warp owns keys m0..m0+15 (= rows iLo / iHi).  dV = P^T dO, then dK = dS^T Q ----
#pragma unroll 1
  for(int pass = 0; pass < 2; ++pass) {
    const float* sA = pass == 0 ? sP : sdS;
    const float* sB = pass == 0 ? sdO : sQ;
    if(pass == 1)
      __syncthreads();  // barrier B: the dS tile is complete
#pragma unroll
    for(int n = 0; n < 8; ++n)
      o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    const int aE = iLo ^ (t << 3), aO = aE ^ 4;  // columns m0+gq (and +8) of rows 8s+2t / 8s+2t+1
#pragma unroll 2
    for(int s8 = 0; s8 < nqs * 8; s8 += 8) {
      const float* pa = sA + (s8 + 2 * t) * 64;
      float af[4] = {pa[aE], pa[aE ^ 8], pa[64 + aO], pa[64 + (aO ^ 8)]};
      uint32_t ahi[4] = {toTf32(af[0]), toTf32(af[1]), toTf32(af[2]), toTf32(af[3])};
      const float* bp = sB + (s8 + 2 * t) * 64;
#pragma unroll
      for(int n = 0; n < 8; ++n)
        mmaSplit<X3>(o[n], af, ahi, bp[(n * 8) ^ cE], bp[64 + ((n * 8) ^ cO)]);
    }
    float* dst = pass == 0 ? dv : dk_;
    __nv_bfloat16* dstS = pass == 0 ? dvS : dkS;
    storeStrip((dst && iLo < g.Tk) ? dst + offK + (size_t)iLo * d : nullptr, (dst && iHi < g.Tk) ? dst + offK + (size_t)iHi * d : nullptr, o, 8, t, (pass == 0 ? accV : accK) != 0,
               (dstS && iLo < g.Tk) ? dstS + offK + (size_t)iLo * d : nullptr, (dstS && iHi < g.Tk) ? dstS + offK + (size_t)iHi * d : nullptr);
  }
}

// ------------------------------------------------------------------------------------------
// forward, warp-private, dk == 64 and Tk <= 64: swizzled tiles, P never leaves the registers
// ------------------------------------------------------------------------------------------
// Same strip ownership as gAttentionForwardWarp; additionally the P strip is used as the A operand
// of O = P V straight from the accumulator registers (permuted contraction slots, see the backward
// kernel), the pitch is the compile-time constant 64 (fragment addresses fold into the LDS
// immediates) and the key mask row is fetched once per lane while the tiles are in flight.
template <bool X3>
__global__ void __launch_bounds__(128) gAttentionForwardWarp64(float* __restrict__ out,
                                                               float* __restrict__ probs,
                                                               const float* __restrict__ q,
                                                               const float* __restrict__ k,
                                                               const float* __restrict__ v,
                                                               const float* __restrict__ mask,
                                                               AttnGeom g,
                                                               int Rq,
                                                               int Rk,
                                                               __nv_bfloat16* __restrict__ outS) {
  extern __shared__ __align__(16) float smemF[];
  pdlEnter();
  float* sQ = smemF;          // [Rq][64]; A rows beyond Rq read into sK: finite, never stored
  float* sK = sQ + Rq * 64;   // [Rk][64]
  float* sV = sK + Rk * 64;   // [Rk][64]

  const int bh = blockIdx.x, b = bh / g.H, h = bh - b * g.H;
  const int row0 = blockIdx.y * 64;
  const int d = g.H * 64;
  const int rowsHere = min(64, g.Tq - row0);
  loadHeadSwizzled(sQ, q + ((size_t)b * g.Tq + row0) * d + h * 64, rowsHere, Rq, d);
  loadHeadSwizzled(sK, k + ((size_t)b * g.Tk) * d + h * 64, g.Tk, Rk, d);
  loadHeadSwizzled(sV, v + ((size_t)b * g.Tk) * d + h * 64, g.Tk, Rk, d);
  cpAsyncCommit();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  const int nts = (g.Tk + 7) >> 3;
  const int iLo = row0 + m0 + gq, iHi = iLo + 8;  // query rows of this lane (global index)
  const bool pair = (g.Tk & 1) == 0;

  // additive mask values of the lane's columns (rows lo / hi differ only for per-query masks)
  float mk[8][4];
  {
    const float* mb = mask ? mask + (size_t)b * g.maskRows * g.Tk : nullptr;
    const float* mLo = mb ? mb + (g.maskRows > 1 ? (size_t)min(iLo, g.Tq - 1) * g.Tk : 0) : nullptr;
    const float* mHi = mb ? mb + (g.maskRows > 1 ? (size_t)min(iHi, g.Tq - 1) * g.Tk : 0) : nullptr;
#pragma unroll
    for(int i = 0; i < 8; ++i) {
      const int j = i * 8 + 2 * t;
      mk[i][0] = mk[i][1] = mk[i][2] = mk[i][3] = 0.f;
      if(mb && j < g.Tk) {
        if(pair) {
          float2 x = *reinterpret_cast<const float2*>(mLo + j);
          mk[i][0] = x.x;
          mk[i][1] = x.y;
          if(g.maskRows > 1)
            x = *reinterpret_cast<const float2*>(mHi + j);
          mk[i][2] = x.x;
          mk[i][3] = x.y;
        } else {
          mk[i][0] = mLo[j];
          mk[i][2] = mHi[j];
          if(j + 1 < g.Tk) {
            mk[i][1] = mLo[j + 1];
            mk[i][3] = mHi[j + 1];
          }
        }
      }
    }
  }
  cpAsyncWaitGroup<0>();
  __syncthreads();
  if(m0 >= rowsHere)
    return;  // the whole strip is padding (no block barrier below)

  // ---- S strip = Q_w K^T ----
  const int cA = t ^ (gq << 2);
  float acc[8][4];
#pragma unroll
  for(int i = 0; i < 8; ++i)
    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  {
    const float* ap = sQ + (m0 + gq) * 64;
    const float* bp0 = sK + gq * 64;
#pragma unroll 2
    for(int k0 = 0; k0 < 64; k0 += 8) {
      const int c0 = k0 ^ cA, c1 = c0 ^ 4;
      float af[4] = {ap[c0], ap[c0 + 8 * 64], ap[c1], ap[c1 + 8 * 64]};
      uint32_t ahi[4] = {toTf32(af[0]), toTf32(af[1]), toTf32(af[2]), toTf32(af[3])};
#pragma unroll
      for(int i = 0; i < 8; ++i)
        if(i < nts)
          mmaSplit<X3>(acc[i], af, ahi, bp0[i * 512 + c0], bp0[i * 512 + c1]);
    }
  }

  // ---- scale, mask, row softmax on the accumulators ----
  float mxLo = -3.0e38f, mxHi = -3.0e38f;
#pragma unroll
  for(int i = 0; i < 8; ++i)
#pragma unroll
    for(int e = 0; e < 2; ++e) {
      const bool real = i * 8 + 2 * t + e < g.Tk;
      acc[i][e] = real ? fmaf(acc[i][e], g.scale, mk[i][e]) : -3.0e38f;
      acc[i][2 + e] = real ? fmaf(acc[i][2 + e], g.scale, mk[i][2 + e]) : -3.0e38f;
      mxLo = fmaxf(mxLo, acc[i][e]);
      mxHi = fmaxf(mxHi, acc[i][2 + e]);
    }
  mxLo = fmaxf(mxLo, __shfl_xor_sync(0xffffffffu, mxLo, 1));
  mxLo = fmaxf(mxLo, __shfl_xor_sync(0xffffffffu, mxLo, 2));
  mxHi = fmaxf(mxHi, __shfl_xor_sync(0xffffffffu, mxHi, 1));
  mxHi = fmaxf(mxHi, __shfl_xor_sync(0xffffffffu, mxHi, 2));
  float sumLo = 0.f, sumHi = 0.f;
#pragma unroll
  for(int i = 0; i < 8; ++i)
#pragma unroll
    for(int e = 0; e < 2; ++e) {
      const bool real = i * 8 + 2 * t + e < g.Tk;
      acc[i][e] = real ? __expf(acc[i][e] - mxLo) : 0.f;
      acc[i][2 + e] = real ? __expf(acc[i][2 + e] - mxHi) : 0.f;
      sumLo += acc[i][e];
      sumHi += acc[i][2 + e];
    }
  sumLo += __shfl_xor_sync(0xffffffffu, sumLo, 1);
  sumLo += __shfl_xor_sync(0xffffffffu, sumLo, 2);
  sumHi += __shfl_xor_sync(0xffffffffu, sumHi, 1);
  sumHi += __shfl_xor_sync(0xffffffffu, sumHi, 2);
  const float invLo = 1.f / sumLo, invHi = 1.f / sumHi;
  float* pLo = (probs && iLo < g.Tq) ? probs + (((size_t)b * g.H + h) * g.Tq + iLo) * g.Tk : nullptr;
  float* pHi = (probs && iHi < g.Tq) ? probs + (((size_t)b * g.H + h) * g.Tq + iHi) * g.Tk : nullptr;
#pragma unroll
  for(int i = 0; i < 8; ++i) {
    const int j = i * 8 + 2 * t;
    acc[i][0] *= invLo;
    acc[i][1] *= invLo;
    acc[i][2] *= invHi;
    acc[i][3] *= invHi;
    if(j < g.Tk) {
      if(pair) {
        if(pLo)
          *reinterpret_cast<float2*>(pLo + j) = make_float2(acc[i][0], acc[i][1]);
        if(pHi)
          *reinterpret_cast<float2*>(pHi + j) = make_float2(acc[i][2], acc[i][3]);
      } else {
        if(pLo) {
          pLo[j] = acc[i][0];
          if(j + 1 < g.Tk)
            pLo[j + 1] = acc[i][1];
        }
        if(pHi) {
          pHi[j] = acc[i][2];
          if(j + 1 < g.Tk)
            pHi[j + 1] = acc[i][3];
        }
      }
    }
  }

  // ---- O strip = P_w V, A operand from the P registers (slot t <-> key 8i+2t, slot t+4 <-> key 8i+2t+1) ----
  const int cE = gq ^ (t << 3), cO = cE ^ 4;
  float o[8][4];
#pragma unroll
  for(int n = 0; n < 8; ++n)
    o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for(int i = 0; i < 8; ++i)
    if(i < nts) {
      float af[4] = {acc[i][0], acc[i][2], acc[i][1], acc[i][3]};
      uint32_t ahi[4] = {toTf32(af[0]), toTf32(af[1]), toTf32(af[2]), toTf32(af[3])};
      const float* vp = sV + (i * 8 + 2 * t) * 64;
#pragma unroll
      for(int n = 0; n < 8; ++n)
        mmaSplit<X3>(o[n], af, ahi, vp[(n * 8) ^ cE], vp[64 + ((n * 8) ^ cO)]);
    }
  storeStrip(iLo < g.Tq ? out + ((size_t)b * g.Tq + iLo) * d + h * 64 : nullptr, iHi < g.Tq ? out + ((size_t)b * g.Tq + iHi) * d + h * 64 : nullptr, o, 8, t, false,
             (outS && iLo < g.Tq) ? outS + ((size_t)b * g.Tq + iLo) * d + h * 64 : nullptr, (outS && iHi < g.Tq) ? outS + ((size_t)b * g.Tq + iHi) * d + h * 64 : nullptr);
}

size_t forwardSmem(const AttnGeom& g) {
  return ((size_t)(g.Tq + 2 * g.Tk) * (g.dk + 4) + (size_t)g.Tq * (pad4(g.Tk) + 1)) * sizeof(float);
}
size_t backwardSmem(const AttnGeom& g) {
  return ((size_t)(2 * g.Tq + 2 * g.Tk) * (g.dk + 4) + (size_t)g.Tq * (pad4(g.Tk) + 1) + g.Tq) * sizeof(float);
}
constexpr size_t kSmemLimit = 227 * 1024;

AttnGeom geometry(Tensor q, Tensor k, Tensor mask, int heads, float scale) {
  AttnGeom g;
  int d = q->shape()[-1];
  ABORT_IF(d % heads != 0, "attention: model width must be divisible by the number of heads");
  g.H = heads;
  g.dk = d / heads;
  g.Tq = q->shape()[-2];
  g.Tk = k->shape()[-2];
  g.B = (int)(q->shape().elements() / ((size_t)g.Tq * d));
  ABORT_IF((size_t)g.B * g.Tk * d != (size_t)k->shape().elements(), "attention: queries and keys disagree on the batch size");
  g.maskRows = 1;
  if(mask) {
    size_t me = mask->shape().elements();
    if(me == (size_t)g.B * g.Tk)
      g.maskRows = 1;
    else if(me == (size_t)g.B * g.Tq * g.Tk)
      g.maskRows = g.Tq;
    else
      ABORT("attention: mask must hold B*Tk or B*Tq*Tk elements, got", mask->shape().toString());
  }
  g.scale = scale;
  return g;
}

}  // namespace

bool AttentionFusable(int Tq, int Tk, int dimModel, int heads) {
  if(heads <= 0 || dimModel % heads != 0)
    return false;
  AttnGeom g{1, heads, Tq, Tk, dimModel / heads, 1, 1.f};
  return (g.dk % 4 == 0) && backwardSmem(g) <= kSmemLimit;
}

namespace {
template <class Kernel>
void ensureSmem(Kernel kernel, size_t smem, size_t& configured) {
  if(smem > configured) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)48 * 1024)));
    configured = smem;
  }
}
}  // namespace

// exact = true: 3xTF32 products (fp32-grade, for the fp32 / bf16x3 GEMM modes);
// exact = false: plain tf32 operands, matching the precision of the tf32 / bf16 GEMM modes
void MultiHeadAttention(Tensor out, Tensor probs, const Tensor q, const Tensor k, const Tensor v, const Tensor mask, int heads, float scale, bool exact) {
  device::setDevice(out->getDevice());
  out->takeLazyZero();
  AttnGeom g = geometry(q, k, mask, heads, scale);
  {
    // warp-private kernels: dk == 64 and Tk <= 64 (swizzled tiles, register-resident P), else dk <= 64, Tk <= 128
    static const bool noWarp = std::getenv("MRN_ATTENTION_NO_WARP") != nullptr;
    const int TkP = pad16(g.Tk);
    if(!noWarp && g.dk == 64 && g.Tk <= 64) {
      const int Rk = 8 * ((g.Tk + 7) / 8);
      const int Rq = g.Tq >= 64 ? 64 : 8 * ((g.Tq + 7) / 8);
      const size_t smemW = (size_t)(Rq + 2 * Rk) * 64 * sizeof(float);
      static bool configured = false;
      if(!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gAttentionForwardWarp64<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 64 * 64 * (int)sizeof(float)));
        CUDA_CHECK(cudaFuncSetAttribute(gAttentionForwardWarp64<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 64 * 64 * (int)sizeof(float)));
        CUDA_CHECK(cudaFuncSetAttribute(gAttentionForwardWarp64<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_CHECK(cudaFuncSetAttribute(gAttentionForwardWarp64<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = true;
      }
      float* pp = probs ? probs->data() : nullptr;
      const float* mp = mask ? mask->data() : nullptr;
      dim3 grid(g.B * g.H, (g.Tq + 63) / 64);
      __nv_bfloat16* outS = shadow::produce(out);  // the output projection reads the context as a bf16 operand (BF16S mode)
      if(exact)
        launchPdl(gAttentionForwardWarp64<true>, grid, dim3(128), smemW, cudaStreamOfEngine(), out->data(), pp, (const float*)q->data(), (const float*)k->data(), (const float*)v->data(), mp, g, Rq, Rk, outS);
      else
        launchPdl(gAttentionForwardWarp64<false>, grid, dim3(128), smemW, cudaStreamOfEngine(), out->data(), pp, (const float*)q->data(), (const float*)k->data(), (const float*)v->data(), mp, g, Rq, Rk, outS);
      CUDA_LAUNCH_CHECK();
      return;
    }
    if(!noWarp && g.dk % 8 == 0 && g.dk <= 64 && TkP <= 128) {
      const int ldQ = pitchMod32(g.dk > TkP ? g.dk : TkP, 4);
      size_t smemW = ((size_t)64 * ldQ + (size_t)TkP * pitchMod32(g.dk, 4) + (size_t)TkP * pitchMod32(g.dk, 8)) * sizeof(float);
      float* pp = probs ? probs->data() : nullptr;
      const float* mp = mask ? mask->data() : nullptr;
      dim3 grid(g.B * g.H, (g.Tq + 63) / 64);
      static size_t cfg[4] = {0, 0, 0, 0};
#define LAUNCH_WARP(X3, NTS, slot)                                                                                                      \
  do {                                                                                                                                  \
    ensureSmem(gAttentionForwardWarp<X3, NTS>, smemW, cfg[slot]);                                                                        \
    launchPdl(gAttentionForwardWarp<X3, NTS>, grid, dim3(128), smemW, cudaStreamOfEngine(), out->data(), pp, q->data(), k->data(), v->data(), mp, g); \
  } while(0)
      if(TkP <= 64) {
        if(exact)
          LAUNCH_WARP(true, 8, 0);
        else
          LAUNCH_WARP(false, 8, 1);
      } else {
        if(exact)
          LAUNCH_WARP(true, 16, 2);
        else
          LAUNCH_WARP(false, 16, 3);
      }
#undef LAUNCH_WARP
      CUDA_LAUNCH_CHECK();
      return;
    }
  }
  {
    MmaLayoutFwd L(g);
    size_t smemMma = L.floats() * sizeof(float);
    static const bool forceSimt = std::getenv("MRN_ATTENTION_SIMT") != nullptr;
    if(!forceSimt && g.dk % 8 == 0 && smemMma <= kSmemLimit) {
      static size_t cfgExact = 0, cfgFast = 0;
      float* pp = probs ? probs->data() : nullptr;
      const float* mp = mask ? mask->data() : nullptr;
      if(exact) {
        ensureSmem(gAttentionForwardMma<true>, smemMma, cfgExact);
        launchPdl(gAttentionForwardMma<true>, dim3(g.B * g.H), dim3(kAttnThreads), smemMma, cudaStreamOfEngine(), out->data(), pp, q->data(), k->data(), v->data(), mp, g);
      } else {
        ensureSmem(gAttentionForwardMma<false>, smemMma, cfgFast);
        launchPdl(gAttentionForwardMma<false>, dim3(g.B * g.H), dim3(kAttnThreads), smemMma, cudaStreamOfEngine(), out->data(), pp, q->data(), k->data(), v->data(), mp, g);
      }
      CUDA_LAUNCH_CHECK();
      return;
    }
  }
  size_t smem = forwardSmem(g);
  ABORT_IF(g.dk % 4 != 0 || smem > kSmemLimit, "attention: shape not supported by the fused kernel", g.Tq, g.Tk, g.dk);
  static size_t configured = 0;
  if(smem > configured) {
    CUDA_CHECK(cudaFuncSetAttribute(gAttentionForward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)48 * 1024)));
    configured = smem;
  }
  gAttentionForward<<<g.B * g.H, kAttnThreads, smem, cudaStreamOfEngine()>>>(
      out->data(), probs ? probs->data() : nullptr, q->data(), k->data(), v->data(), mask ? mask->data() : nullptr, g);
  CUDA_LAUNCH_CHECK();
}

void MultiHeadAttentionGrad(Tensor dq, Tensor dk, Tensor dv, const Tensor adj, const Tensor out, const Tensor probs, const Tensor q, const Tensor k, const Tensor v, int heads, float scale, bool exact) {
  device::setDevice(adj->getDevice());
  AttnGeom g = geometry(q, k, nullptr, heads, scale);
  static const bool noWarp = std::getenv("MRN_ATTENTION_NO_WARP") != nullptr;
  if(!noWarp && g.dk == 64 && g.Tq <= 64 && g.Tk <= 64) {
    // first writer assigns; tensors that alias (q, k, v from the same node) are written in the order
    // dQ, dV, dK by the SAME thread of the kernel (same row -> same warp and lane), so the later
    // ones accumulate in program order
    bool accQ = !dq->takeLazyZero();
    bool accV = !dv->takeLazyZero();
    bool accK = !dk->takeLazyZero();
    const int R = 8 * ((std::max(g.Tq, g.Tk) + 7) / 8);  // tile rows: 56 -> 4 x 14 KB = 56 KB, four CTAs per SM
    const size_t smemW = (size_t)4 * R * 64 * sizeof(float);
    static bool configured = false;
    if(!configured) {
      CUDA_CHECK(cudaFuncSetAttribute(gAttentionBackwardWarp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 64 * 64 * (int)sizeof(float)));
      CUDA_CHECK(cudaFuncSetAttribute(gAttentionBackwardWarp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 64 * 64 * (int)sizeof(float)));
      CUDA_CHECK(cudaFuncSetAttribute(gAttentionBackwardWarp<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CUDA_CHECK(cudaFuncSetAttribute(gAttentionBackwardWarp<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      configured = true;
    }
    // adjoints of the q / k / v projections with this kernel as their one writer: leave the bf16 shadows the
    // input- and weight-gradient products read (BF16S mode); never for aliased or accumulated-into tensors
    const bool distinct = dq->rawData() != dk->rawData() && dq->rawData() != dv->rawData() && dk->rawData() != dv->rawData();
    __nv_bfloat16* dqS = (!accQ && distinct) ? shadow::produce(dq) : nullptr;
    __nv_bfloat16* dkS = (!accK && distinct) ? shadow::produce(dk) : nullptr;
    __nv_bfloat16* dvS = (!accV && distinct) ? shadow::produce(dv) : nullptr;
    // shadow-only adjoints (all their readers are products): the fp32 strips are not written at all
    float* dqP = shadow::fp32Target(dq, dqS);
    float* dkP = shadow::fp32Target(dk, dkS);
    float* dvP = shadow::fp32Target(dv, dvS);
    if(exact)
      launchPdl(gAttentionBackwardWarp<true>, dim3(g.B * g.H), dim3(128), smemW, cudaStreamOfEngine(), dqP, dkP, dvP, (const float*)adj->data(), (const float*)probs->data(),
                (const float*)q->data(), (const float*)k->data(), (const float*)v->data(), g, R, (int)accQ, (int)accK, (int)accV, dqS, dkS, dvS);
    else
      launchPdl(gAttentionBackwardWarp<false>, dim3(g.B * g.H), dim3(128), smemW, cudaStreamOfEngine(), dqP, dkP, dvP, (const float*)adj->data(), (const float*)probs->data(),
                (const float*)q->data(), (const float*)k->data(), (const float*)v->data(), g, R, (int)accQ, (int)accK, (int)accV, dqS, dkS, dvS);
    CUDA_LAUNCH_CHECK();
    return;
  }
  // first writer assigns; tensors that alias (k and v from the same node) are written in the
  // order dV, dQ, dK inside the kernel, separated by block barriers, so the later one accumulates
  bool accV = !dv->takeLazyZero();
  bool accQ = !dq->takeLazyZero();
  bool accK = !dk->takeLazyZero();
  {
    MmaLayoutBwd L(g);
    size_t smemMma = L.floats() * sizeof(float);
    static const bool forceSimt = std::getenv("MRN_ATTENTION_SIMT") != nullptr;
    if(!forceSimt && g.dk % 8 == 0 && smemMma <= kSmemLimit && (size_t)g.Tq * g.dk <= (size_t)L.TkP * L.ldT) {
      static size_t cfgExact = 0, cfgFast = 0;
      if(exact) {
        ensureSmem(gAttentionBackwardMma<true>, smemMma, cfgExact);
        launchPdl(gAttentionBackwardMma<true>, dim3(g.B * g.H), dim3(kAttnThreads), smemMma, cudaStreamOfEngine(), dq->data(), dk->data(), dv->data(), (const float*)adj->data(),
                  (const float*)out->data(), (const float*)probs->data(), (const float*)q->data(), (const float*)k->data(), (const float*)v->data(), g, (int)accQ, (int)accK, (int)accV);
      } else {
        ensureSmem(gAttentionBackwardMma<false>, smemMma, cfgFast);
        launchPdl(gAttentionBackwardMma<false>, dim3(g.B * g.H), dim3(kAttnThreads), smemMma, cudaStreamOfEngine(), dq->data(), dk->data(), dv->data(), (const float*)adj->data(),
                  (const float*)out->data(), (const float*)probs->data(), (const float*)q->data(), (const float*)k->data(), (const float*)v->data(), g, (int)accQ, (int)accK, (int)accV);
      }
      CUDA_LAUNCH_CHECK();
      return;
    }
  }
  size_t smem = backwardSmem(g);
  ABORT_IF(g.dk % 4 != 0 || smem > kSmemLimit, "attention backward: shape not supported by the fused kernel", g.Tq, g.Tk, g.dk);
  static size_t configured = 0;
  if(smem > configured) {
    CUDA_CHECK(cudaFuncSetAttribute(gAttentionBackward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)48 * 1024)));
    configured = smem;
  }
  gAttentionBackward<<<g.B * g.H, kAttnThreads, smem, cudaStreamOfEngine()>>>(
      dq->data(), dk->data(), dv->data(), adj->data(), out->data(), probs->data(), q->data(), k->data(), v->data(), g, accQ, accK, accV);
  CUDA_LAUNCH_CHECK();
}

}  // namespace marian
