// Sharded gradient exchange over NVLink peer memory, fused with the optimizer step.
//
// Reference: SyncGraphGroup::execute (src/training/graph_group_sync.cu:125-151): for every shard
// its owner copies the shard of every other GPU's gradient into a scratch tensor (blocking
// cudaMemcpy), adds it with an Element kernel, runs the optimizer on the shard and copies the
// updated parameters back to every GPU - 2N blocking peer copies + N adds per shard, serialised
// on a host thread pool.
//
// Here, one process per GPU.  Every rank maps the parameter and gradient arenas of all ranks
// (CUDA IPC) and an update is three kernels on the engine stream, no host involvement:
//
//   gPeerBarrier      all ranks have finished backward (flags in peer memory)
//   gGatherReduce     shard_sum[i] = sum_r grads_r[shard_offset + i]   (peer LOADS over NVLink),
//                     sum of squares of the shard for the clipping norm in the same pass
//   gAdam(+peers)     clip factor, 1/N, Adam moments, new parameters written to the local arena
//                     AND to the same range of every peer's arena (peer STORES over NVLink)
//   gPeerBarrier      all parameter shards have landed everywhere
//
// i.e. reduce-scatter + optimizer + all-gather without intermediate copies: the owner reads each
// remote gradient element once and writes each parameter element once per peer.  Per rank and
// step: (N-1)/N * P * 4 bytes in each direction over NVLink, 4 * shard streams of HBM traffic.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "kernels/cuda_helpers.h"
#include "kernels/tensor_operators.h"

namespace marian {

namespace {

// Epoch barrier across the ranks of one node.  pad[r] of rank q = "rank r has reached epoch e".
// Thread t signals peer t, then waits for peer t's signal in the own pad.
__global__ void gPeerBarrier(PeerTable pads, int rank, int nranks, int epoch) {
  int t = threadIdx.x;
  if(t < nranks) {
    __threadfence_system();  // everything this GPU wrote before is visible to the peers first
    volatile int* remote = reinterpret_cast<volatile int*>(pads.ptr[t]) + rank;
    *remote = epoch;
    volatile int* own = reinterpret_cast<volatile int*>(pads.ptr[rank]) + t;
    while(*own < epoch) {
    }
    __threadfence_system();
  }
}

// out[i] = sum over ranks of grads_r[offset + i]; *normSq += sum out[i]^2
__global__ void __launch_bounds__(256) gGatherReduce(float* __restrict__ out, float* __restrict__ normSq, PeerTable grads, int nranks, size_t offset, size_t n) {
  __shared__ float smem[32];
  float sq = 0.f;
  size_t n4 = n >> 2;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for(int r = 0; r < nranks; ++r) {
      float4 g = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(grads.ptr[r]) + offset) + i);
      acc.x += g.x;
      acc.y += g.y;
      acc.z += g.z;
      acc.w += g.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
    sq += (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
  }
  sq = blockSum(sq, smem);
  if(threadIdx.x == 0)
    atomicAdd(normSq, sq);
}

// ---- piece-wise exchange: the same reduce-scatter + Adam + all-gather, cut into phases that can run while the
//      backward sweep is still producing the gradients of the other phase (training/graph_group.h) ----------------
// blockIdx.y = piece.  Grids are kept small (the kernels are NVLink bound and share the GPU with the backward sweep).
__global__ void __launch_bounds__(256) gGatherReducePieces(float* __restrict__ sums, float* __restrict__ partialSq, PeerTable grads, int nranks, PieceList pl) {
  __shared__ float smem[32];
  const int k = blockIdx.y;
  const size_t off = pl.off[k];
  float* out = sums + pl.state[k];
  float sq = 0.f;
  const size_t n4 = pl.len >> 2;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    // all peer loads of an element are in flight together: few blocks must cover the NVLink latency (~2 us)
    float4 g[8];
#pragma unroll
    for(int r = 0; r < 8; ++r)
      g[r] = r < nranks ? __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(grads.ptr[r]) + off) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = g[0];
#pragma unroll
    for(int r = 1; r < 8; ++r) {  // rank order, as the collective sum of the other exchange path
      acc.x += g[r].x;
      acc.y += g[r].y;
      acc.z += g[r].z;
      acc.w += g[r].w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
    sq += (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
  }
  sq = blockSum(sq, smem);
  if(threadIdx.x == 0)
    atomicAdd(partialSq + pl.shard[k], sq);
}

__global__ void gPublishPartials(const float* __restrict__ partialSq, PeerTable pads, int rank, int nranks, int phase) {
  // thread (peer, shard): slot [phase][rank][shard] of the peer's pad
  const int peer = threadIdx.x >> 3, s = threadIdx.x & 7;
  if(peer < nranks && s < nranks) {
    volatile float* slot = reinterpret_cast<volatile float*>(reinterpret_cast<uint8_t*>(pads.ptr[peer]) + 1024) + (phase * 8 + rank) * 8 + s;
    *slot = partialSq[s];
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(256) gAdamPieces(PeerTable params, const float* __restrict__ ownPad, int rank, int nranks, int phase, const float* __restrict__ sums, float* __restrict__ m,
                                                   float* __restrict__ v, AdamArgs a, PieceList pl) {
  const int k = blockIdx.y;
  // norm of the whole reference shard: partial sums of squares of all its pieces, published by their owners
  float scale = a.gradScale;
  if(a.clipNorm > 0.f) {
    const float* part = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(ownPad) + 1024) + (size_t)phase * 64 + pl.shard[k];
    float normSq = 0.f;
    for(int r = 0; r < nranks; ++r)
      normSq += part[r * 8];
    const float norm = sqrtf(normSq) * fabsf(a.gradScale);
    if(norm >= a.clipNorm)
      scale *= a.clipNorm / norm;
  }
  const size_t off = pl.off[k], so = pl.state[k];
  float* p = reinterpret_cast<float*>(params.ptr[rank]) + off;
  const size_t n4 = pl.len >> 2;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(sums + so)[i];
    float4 mm = reinterpret_cast<float4*>(m + so)[i];
    float4 vv = reinterpret_cast<float4*>(v + so)[i];
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
    reinterpret_cast<float4*>(m + so)[i] = mm;
    reinterpret_cast<float4*>(v + so)[i] = vv;
    for(int r = 0; r < nranks; ++r)  // all-gather by peer stores
      if(r != rank)
        reinterpret_cast<float4*>(reinterpret_cast<float*>(params.ptr[r]) + off)[i] = pp;
  }
}

// ---- asynchronous sharded parameter server (AsyncGraphGroup) -------------------------------
// A master block lives on its owner GPU:  [int lock | int steps | pad to 256 B][p | m | v], each
// of `shard` floats.  Any rank may fetch a shard or push a gradient slice into it at any time;
// a per-shard spin lock in the owner's memory (system-scope atomics over NVLink) replaces the
// reference's std::mutex shardSync_[idx] (src/training/graph_group_async.cu:16-71).
__global__ void gShardLock(int* lock, int* steps, int* stepsOut) {
  if(threadIdx.x == 0) {
    while(atomicCAS_system(lock, 0, 1) != 0) {
    }
    __threadfence_system();
    if(steps) {  // a push: this update's Adam step number
      int t = *reinterpret_cast<volatile int*>(steps) + 1;
      *reinterpret_cast<volatile int*>(steps) = t;
      *stepsOut = t;
    }
  }
}
__global__ void gShardUnlock(int* lock) {
  if(threadIdx.x == 0) {
    __threadfence_system();
    atomicExch_system(lock, 0);
  }
}

// Adam on a REMOTE master shard with a LOCAL gradient slice: p, m, v are peer memory (read and
// written over NVLink), g and the clipping norm are local.  Formula of optimizers.cu:43-73 with
// the shard's own step counter for the bias correction.
__global__ void __launch_bounds__(256) gAdamRemote(float* p, float* m, float* v, const float* __restrict__ g, size_t n, AdamArgs a, const int* __restrict__ steps, const float* __restrict__ normSq) {
  float scale = a.gradScale;
  if(normSq && a.clipNorm > 0.f) {
    float norm = sqrtf(*normSq) * fabsf(a.gradScale);
    if(norm >= a.clipNorm)
      scale *= a.clipNorm / norm;
  }
  // the step number is only known on the device (it is counted under the shard lock): bias correction in double, as
  // Adam::updateImpl computes it on the host (a float powf drifts from the singleton's results at large t)
  const double t = (double)*steps;
  const float denom1 = (float)(1.0 - pow((double)a.beta1, t)), denom2 = (float)(1.0 - pow((double)a.beta2, t));
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
      P[e] = P[e] - a.eta * (M[e] / denom1) / (sqrtf(V[e] / denom2) + a.eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

}  // namespace

void ShardLock(void* masterBlock, bool countStep, int* stepsOut) {
  int* hdr = reinterpret_cast<int*>(masterBlock);
  gShardLock<<<1, 32, 0, cudaStreamOfEngine()>>>(hdr, countStep ? hdr + 1 : nullptr, stepsOut);
  CUDA_LAUNCH_CHECK();
}
void ShardUnlock(void* masterBlock) {
  gShardUnlock<<<1, 32, 0, cudaStreamOfEngine()>>>(reinterpret_cast<int*>(masterBlock));
  CUDA_LAUNCH_CHECK();
}
void AdamUpdateRemote(void* masterBlock, size_t shardElements, const float* gradSlice, const AdamArgs& args, const int* steps, Tensor normSq) {
  ABORT_IF(shardElements % 4 != 0 || (((uintptr_t)gradSlice) & 15) != 0, "AdamUpdateRemote expects 16-byte aligned shards");
  float* p = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(masterBlock) + 256);
  float* m = p + shardElements;
  float* v = m + shardElements;
  int grid = std::max(1, std::min((int)((shardElements / 4 + 255) / 256), kNumSMs * 8));
  gAdamRemote<<<grid, 256, 0, cudaStreamOfEngine()>>>(p, m, v, gradSlice, shardElements, args, steps, normSq ? normSq->data() : nullptr);
  CUDA_LAUNCH_CHECK();
}

namespace {
// blocks over all pieces of a phase: the phase that runs next to the backward sweep gets about two per SM (enough loads in
// flight for NVLink / HBM, most of the machine left to the sweep), the exposed phase the whole machine
inline dim3 pieceGrid(const PieceList& pl, bool background) {
  static const int bgBudget = std::getenv("MRN_EXCHANGE_BLOCKS") ? std::atoi(std::getenv("MRN_EXCHANGE_BLOCKS")) : 2 * kNumSMs;
  const int budget = background ? bgBudget : 8 * kNumSMs;
  int gx = std::max(1, std::min((int)((pl.len / 4 + 255) / 256), std::max(4, budget / std::max(1, pl.count))));
  return dim3(gx, pl.count);
}
}  // namespace

void PeerGatherReducePieces(Tensor sums, float* partialSq, const PeerTable& grads, int nranks, const PieceList& pl, bool background) {
  device::setDevice(sums->getDevice());
  if(pl.count == 0)
    return;
  ABORT_IF(pl.len % 4 != 0, "peer exchange expects 16-byte aligned pieces");
  gGatherReducePieces<<<pieceGrid(pl, background), 256, 0, cudaStreamOfEngine()>>>(sums->data(), partialSq, grads, nranks, pl);
  CUDA_LAUNCH_CHECK();
}
void PeerPublishPartials(const float* partialSq, const PeerTable& pads, int rank, int nranks, int phase) {
  gPublishPartials<<<1, 64, 0, cudaStreamOfEngine()>>>(partialSq, pads, rank, nranks, phase);
  CUDA_LAUNCH_CHECK();
}
void AdamUpdatePieces(const PeerTable& params, void* ownPad, int rank, int nranks, int phase, Tensor sums, Tensor mt, Tensor vt, const AdamArgs& args, const PieceList& pl, bool background) {
  device::setDevice(sums->getDevice());
  if(pl.count == 0)
    return;
  gAdamPieces<<<pieceGrid(pl, background), 256, 0, cudaStreamOfEngine()>>>(params, (const float*)ownPad, rank, nranks, phase, sums->data(), mt->data(), vt->data(), args, pl);
  CUDA_LAUNCH_CHECK();
}

namespace {
__global__ void gTimeStamp(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  *slot = t;
}
}  // namespace
void DeviceTimeStamp(unsigned long long* slot) {
  gTimeStamp<<<1, 1, 0, cudaStreamOfEngine()>>>(slot);
  CUDA_LAUNCH_CHECK();
}

void PeerBarrier(const PeerTable& pads, int rank, int nranks, int epoch) {
  gPeerBarrier<<<1, 32, 0, cudaStreamOfEngine()>>>(pads, rank, nranks, epoch);
  CUDA_LAUNCH_CHECK();
}

void PeerGatherReduce(Tensor shardSum, Tensor normSq, const PeerTable& grads, int nranks, size_t offset) {
  device::setDevice(shardSum->getDevice());
  size_t n = shardSum->size();
  ABORT_IF(n % 4 != 0 || offset % 4 != 0, "peer exchange expects 16-byte aligned shards");
  normSq->set(0);
  int grid = std::max(1, std::min((int)((n / 4 + 255) / 256), kNumSMs * 8));
  gGatherReduce<<<grid, 256, 0, cudaStreamOfEngine()>>>(shardSum->data(), normSq->data(), grads, nranks, offset, n);
  CUDA_LAUNCH_CHECK();
}

}  // namespace marian
