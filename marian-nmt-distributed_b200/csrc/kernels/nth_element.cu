// Beam search: the n best continuations of every sentence, on the device.
//
// Reference: NthElement::getNBestList (src/translator/nth_element.cu:270-402) - gMaxElement finds the block-wise maxima of
// every sentence's [beam x V] score slab, gMaxElementUpdate extracts one winner and repairs the block maxima, once per
// returned element; the slab it reads was produced by LogSoftmax, a broadcast add of the hypothesis costs and a transpose
// ("make beams continuous", src/translator/beam_search.h:163-176), i.e. the beam x V scores cross HBM four more times.
//
// Here (HBM-bound integer/float streaming work, no tensor cores):
//   gRowCandidates   grid (splits, rows): a CTA keeps its segment of one row in shared memory, leaves (max, sum exp) of the
//                    segment and its n best (value, column) pairs.  Every thread caches the best of its own strided
//                    elements; a round is one block arg-max over the 256 cached pairs and a rescan by the winner only.
//   gMergeCandidates grid (groups): a group is one range (NthElementRanges) or the rows of one sentence
//                    (NthElementLogSoftmax: candidates become prev[row] + ((x - max) - log(sum)), keys become those of the
//                    reference's transposed tensor); n rounds of block arg-max over <= rows*splits*n candidates.
// The fused entry reads the raw logits ONCE: the n best columns of a row do not depend on the row's normaliser, so row
// statistics and candidates come out of the same pass.  Ties go to the lower key in both stages, so the result does not
// depend on the split.  Results are copied to pinned memory and the stream is drained: beam search needs them on the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>

#include "kernels/cuda_helpers.h"
#include "kernels/tensor_operators.h"

namespace marian {
namespace {

constexpr int kThreads = 256;
constexpr int kSegment = 8192;  // floats of a row one CTA keeps in shared memory (32 KB)
constexpr unsigned kNoKey = 0xFFFFFFFFu;

struct Best {
  float v;
  unsigned k;
};

__device__ __forceinline__ bool better(float v, unsigned k, float w, unsigned l) {
  return v > w || (v == w && k < l);
}

// arg-max over the block's (v, k) pairs; every thread returns the winner.  `slot` is 2 * 8 words of shared memory.
__device__ __forceinline__ Best blockBest(float v, unsigned k, Best* slot) {
#pragma unroll
  for(int d = 16; d > 0; d >>= 1) {
    float w = __shfl_xor_sync(0xffffffffu, v, d);
    unsigned l = __shfl_xor_sync(0xffffffffu, k, d);
    if(better(w, l, v, k)) {
      v = w;
      k = l;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // the previous round's readers are done with slot
  if(lane == 0)
    slot[warp] = Best{v, k};
  __syncthreads();
  Best b = slot[0];
#pragma unroll
  for(int w = 1; w < kThreads / 32; ++w) {
    Best o = slot[w];
    if(better(o.v, o.k, b.v, b.k))
      b = o;
  }
  return b;
}

// rows: row r covers the flat elements [rowFirst[r], rowFirst[r+1]) (rowFirst == nullptr: r*V .. r*V+V).
// Output per (row, split): stats (max, sum exp(x - max)) and n (value, column-in-row) pairs, best first.
__global__ void __launch_bounds__(kThreads) gRowCandidates(const float* __restrict__ x, const int* __restrict__ rowFirst, int V, int n, int suppress,
                                                           float2* __restrict__ stats, float* __restrict__ candV, unsigned* __restrict__ candK) {
  extern __shared__ float seg[];
  __shared__ Best slot[kThreads / 32];
  __shared__ float red[kThreads / 32];
  const int row = blockIdx.y, split = blockIdx.x, splits = gridDim.x;
  const long first = rowFirst ? rowFirst[row] : (long)row * V;
  const int len = rowFirst ? rowFirst[row + 1] - rowFirst[row] : V;
  const int begin = min(len, split * kSegment), end = min(len, begin + kSegment);
  const float* src = x + first;
  const size_t out = ((size_t)row * splits + split);

  // pass 1: segment -> shared memory, per-thread maximum and best candidate
  float vmax = -INFINITY;
  float bv = -INFINITY;
  unsigned bk = kNoKey;
  for(int i = begin + threadIdx.x; i < end; i += kThreads) {
    float v = src[i];
    vmax = fmaxf(vmax, v);
    if(i == suppress)
      v = -FLT_MAX;
    seg[i - begin] = v;
    if(better(v, (unsigned)i, bv, bk)) {
      bv = v;
      bk = (unsigned)i;
    }
  }
  // segment maximum and sum of exponentials (of the real values: the suppressed word still counts in the normaliser)
  for(int d = 16; d > 0; d >>= 1)
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
  if((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = vmax;
  __syncthreads();
  float m = red[0];
  for(int w = 1; w < kThreads / 32; ++w)
    m = fmaxf(m, red[w]);
  float sum = 0.f;
  for(int i = begin + threadIdx.x; i < end; i += kThreads) {
    float v = (i == suppress) ? src[i] : seg[i - begin];
    sum += __expf(v - m);
  }
  for(int d = 16; d > 0; d >>= 1)
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
  __syncthreads();
  if((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if(threadIdx.x == 0) {
    float s = 0.f;
    for(int w = 0; w < kThreads / 32; ++w)
      s += red[w];
    stats[out] = make_float2(m, s);
  }

  // n rounds: block arg-max over the cached per-thread bests; only the winner rescans its elements
  for(int j = 0; j < n; ++j) {
    Best b = blockBest(bv, bk, slot);
    if(threadIdx.x == 0) {
      candV[out * n + j] = b.v;
      candK[out * n + j] = b.k;
    }
    if(b.k != kNoKey && bk == b.k) {
      seg[(int)b.k - begin] = -INFINITY;
      bv = -INFINITY;
      bk = kNoKey;
      for(int i = begin + threadIdx.x; i < end; i += kThreads) {
        float v = seg[i - begin];
        if(v != -INFINITY && better(v, (unsigned)i, bv, bk)) {
          bv = v;
          bk = (unsigned)i;
        }
      }
    }
  }
}

// group g merges the candidates of its rows.
//   fused (prev != nullptr): rows b*dimBatch + g, b < rowsPerGroup; key = (g*rowsPerGroup + b)*V + column,
//                            cost = prev[row] + ((x - max_row) - log(sum_row)), suppressed word: -FLT_MAX
//   ranges:                  row g; key = rowFirst[g] + column, cost = x
// emits cumN[g+1]-cumN[g] pairs at cumN[g] (cumN == nullptr: n pairs at g*n).
__global__ void __launch_bounds__(kThreads) gMergeCandidates(const float2* __restrict__ stats, const float* __restrict__ candV, const unsigned* __restrict__ candK,
                                                             int splits, int n, const float* __restrict__ prev, const int* __restrict__ rowFirst, const int* __restrict__ cumN,
                                                             int dimBatch, int rowsPerGroup, int V, int suppress, float* __restrict__ outV, unsigned* __restrict__ outK) {
  extern __shared__ float smem[];
  __shared__ Best slot[kThreads / 32];
  const int g = blockIdx.x;
  const int perRow = splits * n;
  const int total = rowsPerGroup * perRow;
  float* cost = smem;
  unsigned* key = (unsigned*)(smem + total);
  float* rowMax = smem + 2 * total;
  float* rowLogZ = rowMax + rowsPerGroup;

  if(prev) {
    for(int b = threadIdx.x; b < rowsPerGroup; b += kThreads) {
      const int row = b * dimBatch + g;
      float m = -INFINITY;
      for(int s = 0; s < splits; ++s)
        m = fmaxf(m, stats[(size_t)row * splits + s].x);
      float z = 0.f;
      for(int s = 0; s < splits; ++s) {
        float2 st = stats[(size_t)row * splits + s];
        if(st.y > 0.f)
          z += st.y * __expf(st.x - m);
      }
      rowMax[b] = m;
      rowLogZ[b] = logf(z);
    }
    __syncthreads();
  }
  for(int c = threadIdx.x; c < total; c += kThreads) {
    const int b = c / perRow;
    const int row = prev ? b * dimBatch + g : g;
    const size_t src = (size_t)row * perRow + (c - b * perRow);
    float v = candV[src];
    unsigned k = candK[src];
    if(k == kNoKey) {
      v = -INFINITY;
    } else if(prev) {
      v = ((int)k == suppress) ? -FLT_MAX : prev[row] + ((v - rowMax[b]) - rowLogZ[b]);
      k = (unsigned)(g * rowsPerGroup + b) * (unsigned)V + k;
    } else {
      k = (unsigned)rowFirst[g] + k;
    }
    cost[c] = v;
    key[c] = k;
  }
  __syncthreads();

  const int want = cumN ? cumN[g + 1] - cumN[g] : n;
  const int at = cumN ? cumN[g] : g * n;
  float bv = -INFINITY;
  unsigned bk = kNoKey;
  int bc = -1;
  auto rescan = [&]() {
    bv = -INFINITY;
    bk = kNoKey;
    bc = -1;
    for(int c = threadIdx.x; c < total; c += kThreads)
      if(key[c] != kNoKey && (bc < 0 || better(cost[c], key[c], bv, bk))) {
        bv = cost[c];
        bk = key[c];
        bc = c;
      }
  };
  rescan();
  for(int j = 0; j < want; ++j) {
    Best b = blockBest(bv, bk, slot);
    if(threadIdx.x == 0) {
      outV[at + j] = b.v;
      outK[at + j] = b.k;
    }
    if(b.k != kNoKey && bc >= 0 && bk == b.k) {
      key[bc] = kNoKey;
      rescan();
    }
  }
}

// persistent scratch of the calling thread's device (grown on demand; beam search reuses it every step)
struct Scratch {
  void* dev{nullptr};
  size_t bytes{0};
  int device{-1};
  char* get(size_t need) {
    int d = device::getDevice();
    if(need > bytes || d != device) {
      if(dev)
        device::freeDevice(dev);
      bytes = std::max(need, (size_t)1 << 20);
      dev = device::mallocDevice(bytes);
      device = d;
    }
    return (char*)dev;
  }
};
thread_local Scratch scratch;

size_t align256(size_t b) {
  return (b + 255) / 256 * 256;
}

void run(const float* x, int rows, int V, const std::vector<int>* rowFirst, int maxLen, int n, int groups, int rowsPerGroup, int dimBatch, const std::vector<float>* prev,
         const std::vector<int>* cumN, int suppress, size_t results, std::vector<float>& outCosts, std::vector<unsigned>& outKeys) {
  ABORT_IF(n < 1 || n > 1024, "NthElement: n out of range:", n);
  ABORT_IF(device::capturing(), "NthElement needs a host round trip and cannot be captured into a CUDA graph");
  const int splits = std::max(1, (maxLen + kSegment - 1) / kSegment);
  const size_t perRow = (size_t)splits * n;
  // layout of the scratch allocation
  size_t oStats = 0, oCandV = align256(oStats + rows * (size_t)splits * sizeof(float2)), oCandK = align256(oCandV + rows * perRow * 4),
         oPrev = align256(oCandK + rows * perRow * 4), oFirst = align256(oPrev + (size_t)rows * 4), oCum = align256(oFirst + ((size_t)rows + 1) * 4),
         oOutV = align256(oCum + ((size_t)groups + 1) * 4), oOutK = align256(oOutV + results * 4), end = align256(oOutK + results * 4);
  char* dev = scratch.get(end);
  // host -> device: previous costs / range tables through the pinned bounce buffer (drained at the end of every call)
  const size_t hostBytes = (size_t)rows * 4 + ((size_t)rows + 1) * 4 + ((size_t)groups + 1) * 4 + results * 8;
  char* pinned = (char*)device::pinnedScratch(hostBytes);
  char* hp = pinned;
  if(prev) {
    std::copy(prev->begin(), prev->end(), (float*)hp);
    device::copyH2D(dev + oPrev, hp, prev->size() * 4);
    hp += (size_t)rows * 4;
  }
  if(rowFirst) {
    std::copy(rowFirst->begin(), rowFirst->end(), (int*)hp);
    device::copyH2D(dev + oFirst, hp, rowFirst->size() * 4);
    hp += ((size_t)rows + 1) * 4;
  }
  if(cumN) {
    std::copy(cumN->begin(), cumN->end(), (int*)hp);
    device::copyH2D(dev + oCum, hp, cumN->size() * 4);
    hp += ((size_t)groups + 1) * 4;
  }
  auto stream = cudaStreamOfEngine();
  const int segFloats = std::min(maxLen, kSegment);
  gRowCandidates<<<dim3(splits, rows), kThreads, (size_t)segFloats * 4, stream>>>(x, rowFirst ? (const int*)(dev + oFirst) : nullptr, V, n, suppress, (float2*)(dev + oStats),
                                                                                     (float*)(dev + oCandV), (unsigned*)(dev + oCandK));
  CUDA_LAUNCH_CHECK();
  const size_t mergeSmem = ((size_t)rowsPerGroup * perRow * 2 + 2 * (size_t)rowsPerGroup) * 4;
  ABORT_IF(mergeSmem > 200 * 1024, "NthElement: too many candidates per sentence:", rowsPerGroup * perRow);
  if(mergeSmem > 48 * 1024)
    CUDA_CHECK(cudaFuncSetAttribute(gMergeCandidates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mergeSmem));
  gMergeCandidates<<<groups, kThreads, mergeSmem, stream>>>((const float2*)(dev + oStats), (const float*)(dev + oCandV), (const unsigned*)(dev + oCandK), splits, n,
                                                            prev ? (const float*)(dev + oPrev) : nullptr, rowFirst ? (const int*)(dev + oFirst) : nullptr,
                                                            cumN ? (const int*)(dev + oCum) : nullptr, dimBatch, rowsPerGroup, V, suppress, (float*)(dev + oOutV),
                                                            (unsigned*)(dev + oOutK));
  CUDA_LAUNCH_CHECK();
  float* hostV = (float*)hp;
  unsigned* hostK = (unsigned*)(hp + results * 4);
  device::copyD2H(hostV, dev + oOutV, results * 4);
  device::copyD2H(hostK, dev + oOutK, results * 4);
  device::synchronize();
  outCosts.insert(outCosts.end(), hostV, hostV + results);
  outKeys.insert(outKeys.end(), hostK, hostK + results);
}

}  // namespace

void NthElementRanges(Tensor scores, const std::vector<int>& rangeFirst, const std::vector<int>& cumN, std::vector<float>& outCosts, std::vector<unsigned>& outKeys) {
  const int ranges = (int)rangeFirst.size() - 1;
  ABORT_IF(ranges < 1 || cumN.size() != rangeFirst.size(), "NthElementRanges: malformed range tables");
  ABORT_IF((size_t)rangeFirst.back() > scores->size(), "NthElementRanges: ranges exceed the tensor");
  int maxLen = 0, n = 0;
  for(int i = 0; i < ranges; ++i) {
    maxLen = std::max(maxLen, rangeFirst[i + 1] - rangeFirst[i]);
    n = std::max(n, cumN[i + 1] - cumN[i]);
  }
  run(scores->data(), ranges, 0, &rangeFirst, maxLen, n, ranges, 1, 1, nullptr, &cumN, -1, (size_t)cumN.back(), outCosts, outKeys);
}

void NthElementLogSoftmax(Tensor logits, const std::vector<float>& prevCosts, int dimBatch, int beam, int n, bool first, int suppressWord, std::vector<float>& outCosts,
                          std::vector<unsigned>& outKeys) {
  const int V = logits->shape()[-1];
  const int rowsPerSentence = first ? 1 : beam;
  const int rows = rowsPerSentence * dimBatch;
  ABORT_IF((size_t)rows * V > logits->size(), "NthElementLogSoftmax: logits smaller than [beam, batch, V]");
  ABORT_IF((int)prevCosts.size() < rows, "NthElementLogSoftmax: one previous cost per competing row is required");
  std::vector<float> prev(prevCosts.begin(), prevCosts.begin() + rows);
  run(logits->data(), rows, V, nullptr, V, n, dimBatch, rowsPerSentence, dimBatch, &prev, nullptr, suppressWord, (size_t)dimBatch * n, outCosts, outKeys);
}

}  // namespace marian
