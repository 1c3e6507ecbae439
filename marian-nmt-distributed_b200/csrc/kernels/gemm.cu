// Prod / ProdBatched / ProdAffine: the one dense contraction of the hot path.
//
// Reference: src/kernels/tensor_operators.cu:543-654 (cublasSgemm /
// cublasSgemmStridedBatched in fp32, tensor-op math disabled) driven by
// DotNodeOp / AffineNodeOp / DotBatchedNodeOp (src/graph/node_operators_binary.h:13-369).
//
// Arithmetic modes (GemmMode, kernels/tensor_operators.h):
//   TF32    tcgen05.mma kind::tf32 straight on the fp32 tensors: TMA loads the
//           raw row-major operands (rounding fp32 -> tf32 in flight), K-major or
//           MN-major shared-memory descriptors absorb all four transpose cases,
//           so there is NO packing pass and no scratch traffic   [throughput]
//   BF16    tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), operands
//           staged by TMA into 128B-swizzled shared memory          [throughput]
//   BF16X3  same kernel on hi/lo-split operands laid out along K:
//             A' = [A_hi | A_lo | A_hi],  B' = [B_hi | B_hi | B_lo]
//           so one pass over K' = 3K accumulates hi*hi + lo*hi + hi*lo in fp32
//           (~2^-16 relative operand error; used for the 1e-4 parity runs)
//   FP32    tiled SIMT fp32 kernel (exact-mode fallback and debugging aid)
//   BF16S   tcgen05.mma kind::f16 on bf16 SHADOW copies of the fp32 tensors (kernels/shadow.h):
//           written by the producing kernels / the optimizer (or by one flat conversion pass),
//           read by TMA in all four transpose cases like the tf32 path - half the operand bytes
//           L2 -> shared memory and twice the MMA rate of tf32            [throughput, headline]
//
// The PACKED tensor-core kernel (modes BF16 / BF16X3) computes  C[M,N] (+)= alpha * A[M,K] B[N,K]^T (+ bias)
// with BOTH operands K-major.  The four transpose cases of the reference API,
// the fp32 -> bf16 conversion and the hi/lo split are all handled by ONE
// packing pass per operand (read fp32 once, write bf16 once; weights are
// packed once per step and cached, see gemmSetStableRange).
//
// Kernel anatomy (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d -> smem ring (STAGES deep),
//               mbarrier expect_tx / complete_tx
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit
//               releases smem stages and finally signals the epilogue
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers ->
//               alpha/beta/bias -> global (128-bit stores; red.add for split-K)
// 2 CTAs are resident per SM (<= 96 KB smem, 128 TMEM columns each) so one
// tile's epilogue overlaps the other's main loop.  Small-M / large-K products
// (weight gradients: K = 3200 rows of the batch) are split along K across
// gridDim.z and combined with atomic adds.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <unordered_map>
#include <type_traits>
#include <unordered_set>

#include "kernels/cuda_helpers.h"
#include "kernels/shadow.h"
#include "kernels/tensor_operators.h"

namespace marian {

// =============================================================================
// context: mode, scratch arena for packed operands, packed-weight cache
// =============================================================================
struct GemmContext {
  int device{0};
  GemmMode mode{GemmMode::FP32};

  struct Chunk {
    uint8_t* base;
    size_t size;
  };
  std::vector<Chunk> chunks;
  size_t cur{0}, off{0};

  // sources inside [stableLo, stableHi) (the parameter arena) keep their packed
  // copies until the next gemmInvalidateCache()
  const uint8_t* stableLo{nullptr};
  const uint8_t* stableHi{nullptr};
  struct Key {
    const void* p;
    int rows, cols, flags;
    bool operator==(const Key& o) const { return p == o.p && rows == o.rows && cols == o.cols && flags == o.flags; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      return std::hash<const void*>()(k.p) ^ ((size_t)k.rows * 1000003u) ^ ((size_t)k.cols * 7919u) ^ ((size_t)k.flags << 20);
    }
  };
  std::unordered_map<Key, void*, KeyHash> cache;

  // BF16S: bf16 copy of the parameter arena (same element order as [stableLo, stableHi))
  __nv_bfloat16* paramShadow{nullptr};
  size_t paramShadowElems{0};
  bool paramFresh{false};                         // the whole copy matches the fp32 arena
  // bias-gradient column sums queued by products (ProdFlushColumnSums)
  struct PendingSums {
    float* sums;
    const __nv_bfloat16* in;
    int rows, cols;
  };
  std::vector<PendingSums> pendingSums;
  std::unordered_map<const void*, void*> paramConverted;  // tensors converted one by one since the last invalidate (-> lane mark behind the conversion)

  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn encodeTiled{nullptr};

  void* take(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    while(true) {
      if(cur < chunks.size() && off + bytes <= chunks[cur].size) {
        void* p = chunks[cur].base + off;
        off += bytes;
        return p;
      }
      if(cur + 1 < chunks.size()) {
        ++cur;
        off = 0;
        continue;
      }
      size_t sz = std::max(bytes, (size_t)256 << 20);
      Chunk c;
      c.base = (uint8_t*)device::mallocDevice(sz);
      c.size = sz;
      chunks.push_back(c);
      cur = chunks.size() - 1;
      off = 0;
    }
  }
};

GemmHandle createGemmContext(int deviceId) {
  auto c = new GemmContext();
  c->device = deviceId;
  return c;
}
void destroyGemmContext(GemmHandle h) {
  if(!h)
    return;
  for(auto& c : h->chunks)
    device::freeDevice(c.base);
  if(h->paramShadow)
    device::freeDevice(h->paramShadow);
  delete h;
}
void setGemmMode(GemmHandle h, GemmMode m) {
  h->mode = m;
  shadow::setEnabled(m == GemmMode::BF16S);
}
GemmMode getGemmMode(GemmHandle h) {
  return h->mode;
}
// ---- bf16 shadows of activations / adjoints: one bump arena per process (one process per GPU),
//      rewound whenever a new tape starts; a generation counter invalidates what tensors of an
//      older tape may still point to -------------------------------------------------------------
namespace {
struct ShadowPool {
  struct Chunk {
    uint8_t* base;
    size_t size;
  };
  std::vector<Chunk> chunks;
  size_t cur{0}, off{0};
  uint32_t generation{1};
  bool enabled{false};
  void* take(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    while(true) {
      if(cur < chunks.size() && off + bytes <= chunks[cur].size) {
        void* p = chunks[cur].base + off;
        off += bytes;
        return p;
      }
      if(cur + 1 < chunks.size()) {
        ++cur;
        off = 0;
        continue;
      }
      size_t sz = std::max(bytes, (size_t)256 << 20);
      Chunk c;
      c.base = (uint8_t*)device::mallocDevice(sz);
      c.size = sz;
      chunks.push_back(c);
      cur = chunks.size() - 1;
      off = 0;
    }
  }
  void rewind() {
    cur = 0;
    off = 0;
    ++generation;
  }
};
ShadowPool g_shadows;

// fp32 -> bf16, flat; 8 elements per thread where alignment allows
__global__ void __launch_bounds__(256) gToBf16(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, size_t n) {
  const size_t n8 = n >> 3;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0);
    u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2);
    u.w = *reinterpret_cast<uint32_t*>(&p3);
    reinterpret_cast<uint4*>(dst)[i] = u;
  }
  if(blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const size_t i = (n8 << 3) + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// Conversion pass on the MAIN stream: a shadow converted on the side stream could be picked up by a
// later main-stream product without any ordering.  Called from the side stream it steps back, converts,
// and forks again (the side stream then waits for the conversion).
void convertToBf16(__nv_bfloat16* dst, const float* src, size_t n) {
  static const bool trace = std::getenv("MRN_SHADOW_TRACE") != nullptr;  // which operands still arrive without a bf16 copy
  if(trace) {
    static std::unordered_map<size_t, size_t>* hist = nullptr;
    if(!hist) {
      hist = new std::unordered_map<size_t, size_t>();
      atexit([] {
        for(auto& kv : *hist)
          fprintf(stderr, "[shadow-trace] converted by the consumer: %zu elements x %zu times\n", kv.first, kv.second);
      });
    }
    (*hist)[n]++;
  }
  const bool side = device::onSide();
  if(side)
    device::returnFromSide();
  ABORT_IF((((uintptr_t)src) & 15) != 0 || (((uintptr_t)dst) & 15) != 0, "bf16 conversion expects 16-byte aligned buffers");
  gToBf16<<<gridFor((n + 7) / 8, 256), 256, 0, cudaStreamOfEngine()>>>(dst, src, n);
  CUDA_LAUNCH_CHECK();
  if(side)
    device::forkSide();
}
}  // namespace

namespace shadow {
bool enabled() {
  return g_shadows.enabled;
}
void setEnabled(bool on) {
  g_shadows.enabled = on;
}
__nv_bfloat16* produce(const Tensor& t) {
  if(!g_shadows.enabled || !t)
    return nullptr;
  MemoryPiece& m = *t->memory();
  if(!m.shadowWanted || (uint8_t*)t->rawData() != m.data() || (m.size() & 15) != 0)
    return nullptr;
  if(!m.shadow || m.shadowGen != g_shadows.generation) {
    m.shadow = g_shadows.take(m.size() / 2);
    m.shadowGen = g_shadows.generation;
  }
  m.shadowValid = true;
  m.shadowMark = nullptr;
  return (__nv_bfloat16*)m.shadow;
}
namespace {
bool g_skipFp32 = std::getenv("MRN_SHADOW_KEEP_FP32") == nullptr;
}
void setSkipFp32(bool on) {
  g_skipFp32 = on && std::getenv("MRN_SHADOW_KEEP_FP32") == nullptr;
}
float* fp32Target(const Tensor& t, const __nv_bfloat16* producedShadow) {
  MemoryPiece& m = *t->memory();
  if(producedShadow && g_skipFp32 && m.shadowOnly) {
    m.fp32Skipped = true;
    m.lazyZero = false;
    return nullptr;
  }
  return t->data();
}
}  // namespace shadow

void gemmInvalidateCache(GemmHandle h) {
  if(!h->pendingSums.empty()) {
    // only after a backward sweep that did not reach its end (an allocation exception inside ExpressionGraph::fits()):
    // the operand copies the queued sums would read are about to be dropped, and so are that sweep's gradients
    static bool warned = false;
    if(!warned)
      fprintf(stderr, "[marian_b200] %zu queued bias-gradient sums of an unfinished backward sweep dropped\n", h->pendingSums.size());
    warned = true;
    h->pendingSums.clear();
  }
  h->cache.clear();
  h->cur = 0;
  h->off = 0;
  h->paramConverted.clear();
  g_shadows.rewind();
}
void gemmSetStableRange(GemmHandle h, const void* lo, size_t bytes) {
  if(h->stableLo != (const uint8_t*)lo || (size_t)(h->stableHi - h->stableLo) != bytes)
    h->paramFresh = false;
  h->stableLo = (const uint8_t*)lo;
  h->stableHi = h->stableLo + bytes;
  if(h->mode == GemmMode::BF16S && h->paramShadowElems != bytes / sizeof(float)) {
    if(h->paramShadow)
      device::freeDevice(h->paramShadow);
    h->paramShadowElems = bytes / sizeof(float);
    h->paramShadow = (__nv_bfloat16*)device::mallocDevice(h->paramShadowElems * sizeof(__nv_bfloat16) + 1024);
    h->paramFresh = false;
  }
}
void gemmAllowShadowOnly(GemmHandle, bool allow) {
  shadow::setSkipFp32(allow);
}
void* gemmParamShadowFor(GemmHandle h, const Tensor& t) {
  if(h->mode != GemmMode::BF16S || !h->paramShadow || !t)
    return nullptr;
  const uint8_t* p = (const uint8_t*)t->rawData();
  if(p < h->stableLo || p + t->size() * sizeof(float) > h->stableHi)
    return nullptr;
  return h->paramShadow + (p - h->stableLo) / sizeof(float);
}
void gemmParamsUpdated(GemmHandle h, bool shadowWritten) {
  h->paramFresh = shadowWritten && h->mode == GemmMode::BF16S && h->paramShadow != nullptr;
  h->paramConverted.clear();
}
void gemmPrepareStep(GemmHandle h) {
  if(h->mode != GemmMode::BF16S || !h->paramShadow || h->paramFresh || !h->stableLo)
    return;
  convertToBf16(h->paramShadow, (const float*)h->stableLo, h->paramShadowElems);
  h->paramFresh = true;
}

namespace {
// bf16 copy of operand `t` for the BF16S product: the parameter-arena copy, the shadow its producer
// wrote, or a conversion now (kept on the memory piece when `t` covers it, so that the forward value
// converted here is found again by the weight-gradient product of the backward pass).
const __nv_bfloat16* ensureShadow(GemmHandle h, const Tensor& t) {
  {
    MemoryPiece& m0 = *t->memory();
    if(m0.fp32Skipped) {  // only the bf16 copy exists
      ABORT_IF(!(m0.shadowValid && m0.shadow && m0.shadowGen == g_shadows.generation), "shadow-only tensor without a valid bf16 copy");
      return (const __nv_bfloat16*)m0.shadow + (size_t)((const uint8_t*)t->rawData() - m0.data()) / sizeof(float);
    }
  }
  const float* src = t->data();  // materialises a lazily-zero tensor
  const uint8_t* p = (const uint8_t*)src;
  const size_t n = t->size();
  if(h->paramShadow && p >= h->stableLo && p + n * sizeof(float) <= h->stableHi) {
    __nv_bfloat16* dst = h->paramShadow + (p - h->stableLo) / sizeof(float);
    if(!h->paramFresh) {
      auto it = h->paramConverted.find(p);
      if(it == h->paramConverted.end()) {
        convertToBf16(dst, src, n);
        h->paramConverted[p] = device::laneMark();  // (the conversion runs on the lane's stream even when called from the side stream)
      } else {
        device::laneWait(it->second);  // converted by a consumer on another lane (tensors/device.h)
      }
    }
    return dst;
  }
  MemoryPiece& m = *t->memory();
  const size_t off = (size_t)(p - m.data()) / sizeof(float);
  if(m.shadowValid && m.shadow && m.shadowGen == g_shadows.generation) {
    device::laneWait(m.shadowMark);
    return (const __nv_bfloat16*)m.shadow + off;
  }
  const bool whole = off == 0 && n * sizeof(float) == m.size();
  __nv_bfloat16* dst = (__nv_bfloat16*)g_shadows.take(n * sizeof(__nv_bfloat16));
  convertToBf16(dst, src, n);
  if(whole) {
    m.shadow = dst;
    m.shadowGen = g_shadows.generation;
    m.shadowValid = true;
    m.shadowMark = device::laneMark();
  }
  return dst;
}
}  // namespace

// ---- optional per-launch timing of the tensor-core kernel (bench.py roofline) ----
namespace {
unsigned long long* g_stampBuffer = nullptr;
}
// tuning aid: per-CTA timestamps of the next tf32 launches are written to `deviceBuffer`
// (5 x u64 per CTA: start, prologue done, first operands, accumulator complete, end); null disarms
void gemmDebugStamps(unsigned long long* deviceBuffer) {
  g_stampBuffer = deviceBuffer;
}
namespace {
struct GemmProfile {
  bool enabled{false};
  double flops{0};
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
  std::vector<std::string> labels;  // one per event pair: "M,N,K,batches,layout,splits"
  // span mode (MRN_GEMM_SPANS=1): no event nodes; every CTA of a profiled launch folds its
  // %globaltimer start / end into per-launch min / max slots -> pure kernel execution spans
  bool spans{false};
  unsigned long long* spanMin{nullptr};
  unsigned long long* spanMax{nullptr};
  size_t spanCount{0};
  static constexpr size_t kMaxSpans = 8192;
};
GemmProfile g_profile;
}  // namespace

// enable != 0: reset and start recording; enable == 0: stop, synchronise and report.
// Event pairs recorded while a step was being CAPTURED are re-stamped by every replay of that
// graph: the report then covers the last replay (one step), measured inside the graph.
void gemmProfile(int enable, double* ms, double* flops, size_t* launches) {
  if(enable) {
    for(auto& e : g_profile.events) {
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
    g_profile.events.clear();
    g_profile.labels.clear();
    g_profile.flops = 0;
    g_profile.enabled = true;
    g_profile.spans = std::getenv("MRN_GEMM_SPANS") != nullptr;
    g_profile.spanCount = 0;
    if(g_profile.spans && !g_profile.spanMin) {
      g_profile.spanMin = (unsigned long long*)device::mallocDevice(GemmProfile::kMaxSpans * 8);
      g_profile.spanMax = (unsigned long long*)device::mallocDevice(GemmProfile::kMaxSpans * 8);
    }
    *ms = 0;
    *flops = 0;
    *launches = 0;
    return;
  }
  g_profile.enabled = false;
  device::synchronize();
  double total = 0;
  // MRN_GEMM_PROFILE_DUMP=<file>: one CSV line per launch (shape, layout, in-graph microseconds)
  FILE* dump = nullptr;
  if(const char* path = std::getenv("MRN_GEMM_PROFILE_DUMP"))
    dump = fopen((std::string(path) + (g_profile.spans ? ".spans" : "")).c_str(), "w");
  if(g_profile.spans) {
    size_t n = std::min(g_profile.spanCount, GemmProfile::kMaxSpans);
    std::vector<unsigned long long> mn(n), mx(n);
    if(n) {
      CUDA_CHECK(cudaMemcpy(mn.data(), g_profile.spanMin, n * 8, cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemcpy(mx.data(), g_profile.spanMax, n * 8, cudaMemcpyDeviceToHost));
    }
    for(size_t i = 0; i < n; ++i) {
      double us = mx[i] > mn[i] ? (double)(mx[i] - mn[i]) / 1000.0 : 0.0;
      total += us / 1000.0;
      if(dump)
        fprintf(dump, "%s,%.2f\n", i < g_profile.labels.size() ? g_profile.labels[i].c_str() : "?", us);
    }
    if(dump)
      fclose(dump);
    *ms = total;
    *flops = g_profile.flops;
    *launches = n;
    return;
  }
  for(size_t i = 0; i < g_profile.events.size(); ++i) {
    auto& e = g_profile.events[i];
    float t = 0;
    if(cudaEventElapsedTime(&t, e.first, e.second) == cudaSuccess)
      total += t;
    if(dump)
      fprintf(dump, "%s,%.2f\n", i < g_profile.labels.size() ? g_profile.labels[i].c_str() : "?", t * 1000.f);
  }
  if(dump)
    fclose(dump);
  *ms = total;
  *flops = g_profile.flops;
  *launches = g_profile.events.size();
}

namespace {

// Event pair around one tensor-core launch.  Eagerly these are plain records; while the step is
// being captured into a CUDA graph they become EXTERNAL event-record nodes, so every replay of
// the graph re-stamps them and the host can read the in-graph duration of each launch.
struct ProfileScope {
  cudaEvent_t e0{nullptr}, e1{nullptr};
  bool on{false};
  unsigned long long* spanMin{nullptr};
  unsigned long long* spanMax{nullptr};
  explicit ProfileScope(double flops) {
    on = g_profile.enabled;
    if(!on)
      return;
    if(g_profile.spans) {
      if(g_profile.spanCount == 0) {
        // (re)armed at the first profiled launch of a step - inside the captured graph too
        CUDA_CHECK(cudaMemsetAsync(g_profile.spanMin, 0xFF, GemmProfile::kMaxSpans * 8, cudaStreamOfEngine()));
        CUDA_CHECK(cudaMemsetAsync(g_profile.spanMax, 0x00, GemmProfile::kMaxSpans * 8, cudaStreamOfEngine()));
      }
      if(g_profile.spanCount < GemmProfile::kMaxSpans) {
        spanMin = g_profile.spanMin + g_profile.spanCount;
        spanMax = g_profile.spanMax + g_profile.spanCount;
      }
      g_profile.spanCount++;
      g_profile.flops += flops;
      return;
    }
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    record(e0);
    g_profile.flops += flops;
  }
  void finish(const std::string& label = std::string()) {
    if(!on)
      return;
    if(g_profile.spans) {
      g_profile.labels.push_back(label);
      return;
    }
    record(e1);
    g_profile.events.push_back({e0, e1});
    g_profile.labels.push_back(label);
  }
  static void record(cudaEvent_t e) {
    if(device::capturing())
      CUDA_CHECK(cudaEventRecordWithFlags(e, cudaStreamOfEngine(), cudaEventRecordExternal));
    else
      CUDA_CHECK(cudaEventRecord(e, cudaStreamOfEngine()));
  }
};

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int kStagePitch = 36;  // floats per staged epilogue row: 144 B keeps float4 alignment, conflict-free
constexpr int UMMA_K = 16;

// =============================================================================
// operand packing: fp32 [rows, cols] (row-major) -> bf16 K-major [outRows, Kp]
// =============================================================================
enum PackFlags { PACK_TRANSPOSE = 1, PACK_X3 = 2, PACK_ROLE_B = 4 };

struct PackGeom {
  int srcRows, srcCols;  // per batch
  int outRows;           // valid output rows per batch (= srcRows, or srcCols when transposed)
  int outRowsPad;        // rows per batch in the packed buffer (zero padded)
  int K, Kp;             // reduction length and its 64-padded size
  int segs;              // 1, or 3 for the hi/lo split
  int roleB;
  size_t srcBatchStride;
};

__device__ __forceinline__ void splitBf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// segment s of the split layout holds hi or lo depending on the operand role:
//   A: [hi, lo, hi]   B: [hi, hi, lo]
__device__ __forceinline__ bool segmentIsLo(int s, int roleB) {
  return roleB ? (s == 2) : (s == 1);
}

// no transpose: out[r][k] = src[r][k]; one thread per (row, pair of k)
__global__ void __launch_bounds__(256) gPackRows(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, PackGeom g, int batches) {
  int kPairs = g.Kp >> 1;
  long long perBatch = (long long)g.outRowsPad * kPairs;
  long long items = perBatch * batches;
  size_t dstRowElems = (size_t)g.Kp * g.segs;
  for(long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += (long long)gridDim.x * blockDim.x) {
    int b = (int)(w / perBatch);
    long long rem = w - (long long)b * perBatch;
    int r = (int)(rem / kPairs);
    int k = (int)(rem - (long long)r * kPairs) << 1;
    float x0 = 0.f, x1 = 0.f;
    if(r < g.outRows) {
      const float* s = src + (size_t)b * g.srcBatchStride + (size_t)r * g.srcCols;
      if(k < g.K)
        x0 = s[k];
      if(k + 1 < g.K)
        x1 = s[k + 1];
    }
    __nv_bfloat16 h0, l0, h1, l1;
    splitBf16(x0, h0, l0);
    splitBf16(x1, h1, l1);
    __nv_bfloat16* d = dst + ((size_t)b * g.outRowsPad + r) * dstRowElems + k;
    for(int s = 0; s < g.segs; ++s) {
      bool lo = g.segs == 3 && segmentIsLo(s, g.roleB);
      __nv_bfloat162 v;
      v.x = lo ? l0 : h0;
      v.y = lo ? l1 : h1;
      *reinterpret_cast<__nv_bfloat162*>(d + (size_t)s * g.Kp) = v;
    }
  }
}

// transpose: out[r][k] = src[k][r]; 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) gPackTranspose(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, PackGeom g) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  const float* s = src + (size_t)b * g.srcBatchStride;
  int r0 = blockIdx.y * 32;  // output rows  = source columns
  int k0 = blockIdx.x * 32;  // output K     = source rows
  for(int i = threadIdx.y; i < 32; i += blockDim.y) {
    int sr = k0 + i, sc = r0 + threadIdx.x;
    tile[i][threadIdx.x] = (sr < g.srcRows && sc < g.srcCols) ? s[(size_t)sr * g.srcCols + sc] : 0.f;
  }
  __syncthreads();
  size_t dstRowElems = (size_t)g.Kp * g.segs;
  for(int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, k = k0 + threadIdx.x;
    if(r < g.outRowsPad && k < g.Kp) {
      float x = tile[threadIdx.x][i];  // zero outside the source by construction
      __nv_bfloat16 hi, lo;
      splitBf16(x, hi, lo);
      __nv_bfloat16* d = dst + ((size_t)b * g.outRowsPad + r) * dstRowElems + k;
      for(int sgm = 0; sgm < g.segs; ++sgm)
        d[(size_t)sgm * g.Kp] = (g.segs == 3 && segmentIsLo(sgm, g.roleB)) ? lo : hi;
    }
  }
}

inline int roundUp(int x, int m) {
  return (x + m - 1) / m * m;
}

struct Packed {
  __nv_bfloat16* data;
  int rowsPad;  // rows per batch
  int Ktotal;   // Kp * segs
};

// Packs `batches` matrices [srcRows, srcCols] into K-major bf16.
Packed packOperand(GemmHandle h, const float* src, int srcRows, int srcCols, int batches, size_t srcBatchStride, bool transpose, bool roleB, bool x3, int padRowsTo) {
  PackGeom g;
  g.srcRows = srcRows;
  g.srcCols = srcCols;
  g.outRows = transpose ? srcCols : srcRows;
  g.K = transpose ? srcRows : srcCols;
  g.Kp = roundUp(g.K, BLOCK_K);
  g.segs = x3 ? 3 : 1;
  g.roleB = roleB;
  g.outRowsPad = padRowsTo > 0 ? roundUp(g.outRows, padRowsTo) : g.outRows;
  g.srcBatchStride = srcBatchStride;

  int flags = (transpose ? PACK_TRANSPOSE : 0) | (x3 ? PACK_X3 : 0) | (roleB ? PACK_ROLE_B : 0) | (padRowsTo << 4) | (batches << 12);
  GemmContext::Key key{src, srcRows, srcCols, flags};
  bool stable = h->stableLo && (const uint8_t*)src >= h->stableLo && (const uint8_t*)src < h->stableHi;
  if(stable) {
    auto it = h->cache.find(key);
    if(it != h->cache.end())
      return Packed{(__nv_bfloat16*)it->second, g.outRowsPad, g.Kp * g.segs};
  }

  size_t elems = (size_t)batches * g.outRowsPad * g.Kp * g.segs;
  auto dst = (__nv_bfloat16*)h->take(elems * sizeof(__nv_bfloat16));
  auto st = cudaStreamOfEngine();
  if(!transpose) {
    long long items = (long long)batches * g.outRowsPad * (g.Kp / 2);
    gPackRows<<<gridFor((size_t)items, 256), 256, 0, st>>>(dst, src, g, batches);
  } else {
    dim3 grid((g.Kp + 31) / 32, (g.outRowsPad + 31) / 32, batches);
    gPackTranspose<<<grid, dim3(32, 8), 0, st>>>(dst, src, g);
  }
  CUDA_LAUNCH_CHECK();
  if(stable)
    h->cache[key] = dst;
  return Packed{dst, g.outRowsPad, g.Kp * g.segs};
}

// =============================================================================
// fp32 SIMT GEMM (exact mode)
// =============================================================================
struct SimtArgs {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  int M, N, K;
  int lda, ldb, ldc;
  int transA, transB;
  float alpha, beta;
  size_t strideA, strideB, strideC;
};

__global__ void __launch_bounds__(256) gGemmSimt(SimtArgs a) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  int b = blockIdx.z;
  const float* A = a.A + (size_t)b * a.strideA;
  const float* B = a.B + (size_t)b * a.strideB;
  float* C = a.C + (size_t)b * a.strideC;
  int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for(int i = 0; i < 4; ++i)
#pragma unroll
    for(int j = 0; j < 4; ++j)
      acc[i][j] = 0.f;

  for(int k0 = 0; k0 < a.K; k0 += TK) {
    for(int e = threadIdx.x; e < TM * TK; e += 256) {
      // consecutive threads walk the contiguous source dimension
      int mm, kk;
      if(a.transA) {
        mm = e % TM;
        kk = e / TM;
      } else {
        kk = e % TK;
        mm = e / TK;
      }
      int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if(m < a.M && k < a.K)
        v = a.transA ? A[(size_t)k * a.lda + m] : A[(size_t)m * a.lda + k];
      As[kk][mm] = v;
    }
    for(int e = threadIdx.x; e < TN * TK; e += 256) {
      int nn, kk;
      if(a.transB) {
        kk = e % TK;
        nn = e / TK;
      } else {
        nn = e % TN;
        kk = e / TN;
      }
      int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if(n < a.N && k < a.K)
        v = a.transB ? B[(size_t)n * a.ldb + k] : B[(size_t)k * a.ldb + n];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for(int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for(int i = 0; i < 4; ++i)
        av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for(int j = 0; j < 4; ++j)
        bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for(int i = 0; i < 4; ++i)
#pragma unroll
        for(int j = 0; j < 4; ++j)
          acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for(int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if(m >= a.M)
      continue;
#pragma unroll
    for(int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if(n >= a.N)
        continue;
      float v = a.alpha * acc[i][j];
      if(a.bias)
        v += a.bias[n];
      size_t idx = (size_t)m * a.ldc + n;
      if(a.beta != 0.f)
        v += a.beta * C[idx];
      C[idx] = v;
    }
  }
}

// =============================================================================
// tcgen05 / TMA / TMEM kernel
// =============================================================================
__device__ __forceinline__ uint32_t smemAddr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smemAddr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmaLoad2D(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smemAddr(dst)),
      "l"((uint64_t)map),
      "r"(smemAddr(bar)),
      "r"(c0),
      "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmemD),
      "l"(descA),
      "l"(descB),
      "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ummaCommit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void tcgenFenceBefore() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgenFenceAfter() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (= 1, unused for SW128 K-major)
//   bits [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 -> 64)
//   bits [46,48) version = 1 (Blackwell)     bits [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t makeSmemDesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)64 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for bf16 x bf16 -> fp32, both K-major
__host__ __device__ constexpr uint32_t makeInstrDesc(int M, int N) {
  return (1u << 4)     // c_format  = F32
         | (1u << 7)   // a_format  = BF16
         | (1u << 10)  // b_format  = BF16
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcArgs {
  float* C;
  const float* bias;
  int M, N;
  int ldc;
  int kBlocks;          // total 64-wide K blocks
  int kBlocksPerSplit;  // blocks handled by one z-slice
  int kBlocksGroup;     // K-grouped products: k-blocks per (A_g, B_g) pair; kBlocks = groups * kBlocksGroup
  int splits;
  int rowsPerBatchA, rowsPerBatchB;  // row pitch between batches in the packed operands (0 = shared)
  size_t strideC;
  float alpha, beta;
  int atomicOut;  // combine with red.add (split-K)
  float* colSum[3];   // != null: column sums of the (K-major) A operand of group g are red.add-ed here (bias gradients)
  int colSumLen;      // columns of one A operand
  const float* gate;  // GATE kernels: pre-activation h, same layout as C; the product is scaled by swish'(h)
  __nv_bfloat16* shadowC;  // != null: bf16 copy of the final C values (BF16S: C is itself a product operand later)
  int tmaStore;            // bf16 kernels: 1 = epilogue leaves through TMA tensor stores, 2 = through TMA reduce-add (C += tile)
  unsigned long long* stamps;  // tuning aid: per-CTA %globaltimer stamps (5 per CTA), or null
  unsigned long long* spanMin;  // profiling: per-launch min(start) / max(end) over the CTAs, or null
  unsigned long long* spanMax;
};

template <int BN, bool GATE = false>
__device__ __forceinline__ void epilogueTile(const TcArgs& a, uint32_t tmemBase, uint64_t* tmemFullBar, float* stage, int warp, int lane, int m0, int n0, int batch, int split, uint32_t parity = 0, int cBegin = 0, int cEnd = BN);
template <int BN, bool GATE = false>
__device__ __forceinline__ void epilogueTileTma(const TcArgs& a, const CUtensorMap* tmC, uint32_t tmemBase, uint64_t* tmemFullBar, uint8_t* sbuf, int warp, int lane, int m0, int n0, int batch, int split, uint32_t parity = 0);

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) gGemmTcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcArgs a) {
  typedef TcSmem<BN, STAGES> L;
  extern __shared__ uint8_t smemRaw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
  uint64_t* fullBar = (uint64_t*)(smem + L::BAR_OFFSET);
  uint64_t* emptyBar = fullBar + STAGES;
  uint64_t* tmemFullBar = emptyBar + STAGES;
  uint32_t* tmemHolder = (uint32_t*)(tmemFullBar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * BN;
  const int batch = blockIdx.z / a.splits;
  const int split = blockIdx.z - batch * a.splits;
  const int kb0 = split * a.kBlocksPerSplit;
  const int nkb = min(a.kBlocksPerSplit, a.kBlocks - kb0);

  if(warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    for(int s = 0; s < STAGES; ++s) {
      mbarInit(fullBar + s, 1);
      mbarInit(emptyBar + s, 1);
    }
    mbarInit(tmemFullBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if(warp == 1) {
    // 128 lanes x BN fp32 columns of tensor memory for the accumulator
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(tmemHolder)), "r"((uint32_t)(BN < 32 ? 32 : BN)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgenFenceBefore();
  __syncthreads();
  tcgenFenceAfter();
  const uint32_t tmemBase = *tmemHolder;

  if(warp == 0) {
    if(lane == 0) {
      // ---------------- TMA producer ----------------
      const int rowA = batch * a.rowsPerBatchA + m0;
      const int rowB = batch * a.rowsPerBatchB + n0;
      for(int i = 0; i < nkb; ++i) {
        int s = i % STAGES;
        uint32_t phase = (uint32_t)(i / STAGES) & 1u;
        mbarWait(emptyBar + s, phase ^ 1u);
        mbarExpectTx(fullBar + s, (uint32_t)L::STAGE_BYTES);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        int kc = (kb0 + i) * BLOCK_K;
        tmaLoad2D(&tmA, fullBar + s, sa, kc, rowA);
        tmaLoad2D(&tmB, fullBar + s, sb, kc, rowB);
      }
    }
  } else if(warp == 1) {
    if(lane == 0) {
      // ---------------- MMA issuer (single thread) ----------------
      constexpr uint32_t idesc = makeInstrDesc(BLOCK_M, BN);
      for(int i = 0; i < nkb; ++i) {
        int s = i % STAGES;
        uint32_t phase = (uint32_t)(i / STAGES) & 1u;
        mbarWait(fullBar + s, phase);
        tcgenFenceAfter();
        uint32_t sa = smemAddr(smem + s * L::STAGE_BYTES);
        uint32_t sb = sa + L::A_BYTES;
        uint64_t descA = makeSmemDesc(sa);
        uint64_t descB = makeSmemDesc(sb);
#pragma unroll
        for(int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (>>4) address field
          umma(tmemBase, descA + (uint64_t)(k * 2), descB + (uint64_t)(k * 2), idesc, (uint32_t)((i | k) != 0));
        }
        ummaCommit(emptyBar + s);  // frees the smem stage once these MMAs retire
      }
      ummaCommit(tmemFullBar);  // accumulator complete
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> smem staging -> global ----------------
    float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * kStagePitch);
    epilogueTile<BN>(a, tmemBase, tmemFullBar, stage, warp, lane, m0, n0, batch, split);
  }

  tcgenFenceBefore();
  __syncthreads();
  if(warp == 1) {
    __syncwarp();
    tcgenFenceAfter();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)(BN < 32 ? 32 : BN)));
  }
}


// =============================================================================
// tf32 kernel: operands are the fp32 tensors themselves
// =============================================================================
// Shared-memory tile of one operand per stage = 32 reduction elements x ROWS:
//   K-major  (reduction dim contiguous in global memory): ONE TMA box {32 k, ROWS}; rows of
//            128 bytes, SWIZZLE_128B; descriptor SBO = 1024 B (8 rows), k-step (8 tf32 = 32 B)
//            advances the start address inside the swizzle row.
//   MN-major (the M / N dim contiguous: transposed operands, i.e. weights in the forward
//            product and both operands of dW = X^T dY): ROWS/32 TMA boxes {32 mn, 32 k}; each
//            box is 32 k-rows of 128 bytes, swizzled in 32-byte atoms
//            (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> UMMA SWIZZLE_128B_BASE32B, the only
//            MN-major layout the tensor core accepts for 32-bit elements); descriptor
//            LBO = 4096 B (next 32 mn), SBO = 512 B (next 4 k-rows), k-step (8 k-rows) = 1024 B.
constexpr int TF_BLOCK_K = 32;

__device__ __forceinline__ void tmaLoad3D(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smemAddr(dst)),
      "l"((uint64_t)map),
      "r"(smemAddr(bar)),
      "r"(c0),
      "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void ummaTf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmemD),
      "l"(descA),
      "l"(descB),
      "r"(idesc),
      "r"(accumulate)
      : "memory");
}
template <bool MN>
__device__ __forceinline__ uint64_t makeSmemDescTf32(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  if(MN) {
    d |= (uint64_t)(4096 >> 4) << 16;  // LBO: next block of 32 mn
    d |= (uint64_t)(512 >> 4) << 32;   // SBO: next 4 k-rows
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;            // SWIZZLE_128B_BASE32B
  } else {
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;  // SBO: next 8 rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  }
  return d;
}
__host__ __device__ constexpr uint32_t makeInstrDescTf32(int M, int N, bool aMN, bool bMN) {
  return (1u << 4)     // c_format = F32
         | (2u << 7)   // a_format = TF32
         | (2u << 10)  // b_format = TF32
         | ((aMN ? 1u : 0u) << 15) | ((bMN ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN, int STAGES>
struct TfSmem {
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;
};

// Shared epilogue of both tensor-core kernels: TMEM -> registers -> smem staging -> global
// with alpha / beta / bias; `stage` is this warp's 32 x kStagePitch float scratch.
// beta != 0 (gradient accumulation): the old C values of a 32-column block are requested BEFORE
// the accumulator is waited for / read back, so their latency hides behind the main loop's tail
// and the TMEM read of the previous block.
// d swish(h) / dh = s (1 + h (1 - s)), s = sigmoid(h); ex2.approx / rcp.approx like the element-wise functor
__device__ __forceinline__ float swishGrad(float h) {
  float z = __expf(-fabsf(h));
  float r = __fdividef(1.f, 1.f + z);
  float sg = h > 0.f ? r : z * r;
  return sg * (1.f + h * (1.f - sg));
}

// GATE: C = beta C + (alpha acc + bias) o swish'(gate) - the backward pass of "affine after swish"
// (dH = (dY W^T) o swish'(H)) without the intermediate adjoint and its element-wise kernel.
template <int BN, bool GATE>
__device__ __forceinline__ void epilogueTile(const TcArgs& a, uint32_t tmemBase, uint64_t* tmemFullBar, float* stage, int warp, int lane, int m0, int n0, int batch, int split, uint32_t parity, int cBegin, int cEnd) {
  // [cBegin, cEnd): the 32-column blocks of the tile this warp handles (two warps may share a TMEM lane quarter)
  const int q = warp & 3;  // TMEM lane quarter this warp may access
  float* Cb = a.C + (size_t)batch * a.strideC;
  const float* Gb = GATE ? a.gate + (size_t)batch * a.strideC : nullptr;
  const bool addBias = a.bias != nullptr && split == 0;
  const int rowBase = m0 + q * 32;
  const int sub = lane >> 3;      // row within a group of 4
  const int cq = (lane & 7) * 4;  // first of this lane's 4 columns
  const bool readOld = a.beta != 0.f && !a.atomicOut;
  const uint32_t stageAddr = smemAddr(stage);

  auto isVec = [&](int col0) { return col0 + 32 <= a.N && ((a.ldc & 3) == 0) && (((((uintptr_t)(Cb + col0)) & 15) == 0) || !a.C); };
  float4 oldv[8];
  auto prefetchOld = [&](int col0) {
    if(!readOld || col0 >= a.N || rowBase >= a.M || !isVec(col0))
      return;
#pragma unroll
    for(int i = 0; i < 8; ++i) {
      int grow = rowBase + i * 4 + sub;
      oldv[i] = grow < a.M ? *reinterpret_cast<const float4*>(Cb + (size_t)grow * a.ldc + col0 + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  float4 gatev[GATE ? 8 : 1];
  auto prefetchGate = [&](int col0) {
    if(!GATE || col0 >= a.N || rowBase >= a.M || !isVec(col0) || (((uintptr_t)(Gb + col0)) & 15) != 0)
      return;
#pragma unroll
    for(int i = 0; i < (GATE ? 8 : 1); ++i) {
      int grow = rowBase + i * 4 + sub;
      gatev[i] = grow < a.M ? *reinterpret_cast<const float4*>(Gb + (size_t)grow * a.ldc + col0 + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  prefetchOld(n0 + cBegin);
  prefetchGate(n0 + cBegin);
  // bias of this lane's four columns in every 32-column block, requested before the accumulator wait (next to no
  // L1 under the maximal shared-memory carve-out: inside the loop each load is an exposed L2 round trip)
  auto biasOf = [&](int col0) { return (addBias && col0 < a.N && isVec(col0)) ? *reinterpret_cast<const float4*>(a.bias + col0 + cq) : make_float4(0.f, 0.f, 0.f, 0.f); };
  float4 bq = biasOf(n0 + cBegin);
  mbarWait(tmemFullBar, parity);
  tcgenFenceAfter();
#pragma unroll 1
  for(int c0 = cBegin; c0 < cEnd; c0 += 32) {
    const int col0 = n0 + c0;
    if(col0 >= a.N || rowBase >= a.M)
      break;
    uint32_t r[32];
    uint32_t taddr = tmemBase + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    {
      // explicit shared-space accesses: through the generic pointer these compile to generic ST.E / LD.E
      const uint32_t srow = stageAddr + (uint32_t)(lane * kStagePitch * 4);
#pragma unroll
      for(int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + 16u * j), "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
    }
    __syncwarp();
    const int ncols = min(32, a.N - col0);
    if(isVec(col0) && (!GATE || (((uintptr_t)(Gb + col0)) & 15) == 0)) {
      float4 outv[8];
#pragma unroll
      for(int i = 0; i < 8; ++i) {
        float4 acc;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(acc.x), "=f"(acc.y), "=f"(acc.z), "=f"(acc.w) : "r"(stageAddr + (uint32_t)(((i * 4 + sub) * kStagePitch + cq) * 4)));
        float4 v;
        v.x = a.alpha * acc.x + bq.x;
        v.y = a.alpha * acc.y + bq.y;
        v.z = a.alpha * acc.z + bq.z;
        v.w = a.alpha * acc.w + bq.w;
        if(GATE) {
          const float4 gv = gatev[GATE ? i : 0];
          v.x *= swishGrad(gv.x);
          v.y *= swishGrad(gv.y);
          v.z *= swishGrad(gv.z);
          v.w *= swishGrad(gv.w);
        }
        if(readOld) {
          v.x += a.beta * oldv[i].x;
          v.y += a.beta * oldv[i].y;
          v.z += a.beta * oldv[i].z;
          v.w += a.beta * oldv[i].w;
        }
        outv[i] = v;
      }
      prefetchOld(col0 + 32);  // next block's old values travel while this one is stored
      prefetchGate(col0 + 32);
      bq = biasOf(col0 + 32);
#pragma unroll
      for(int i = 0; i < 8; ++i) {
        int grow = rowBase + i * 4 + sub;
        if(grow < a.M) {
          float* cp = Cb + (size_t)grow * a.ldc + col0 + cq;
          if(a.atomicOut) {
            // split-K partial sums: vector reduction straight into L2 (sm_90+)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(outv[i].x), "f"(outv[i].y), "f"(outv[i].z), "f"(outv[i].w) : "memory");
          } else {
            if(a.C)  // (null: shadow-only output, every reader takes the bf16 copy)
              *reinterpret_cast<float4*>(cp) = outv[i];
            shadow::store4(a.shadowC, (size_t)batch * a.strideC + (size_t)grow * a.ldc + col0 + cq, outv[i]);
          }
        }
      }
    } else {
      if(lane < ncols) {
        // general path: lane = column, rows walked one by one (still contiguous per row)
        float bv = addBias ? a.bias[col0 + lane] : 0.f;
        int rmax = min(32, a.M - rowBase);
        for(int rloc = 0; rloc < rmax; ++rloc) {
          float v = a.alpha * stage[rloc * kStagePitch + lane] + bv;
          if(GATE)
            v *= swishGrad(Gb[(size_t)(rowBase + rloc) * a.ldc + col0 + lane]);
          float* cp = Cb + (size_t)(rowBase + rloc) * a.ldc + col0 + lane;
          if(a.atomicOut) {
            atomicAdd(cp, v);
          } else {
            if(a.beta != 0.f)
              v += a.beta * *cp;
            if(a.C)
              *cp = v;
            shadow::store1(a.shadowC, (size_t)batch * a.strideC + (size_t)(rowBase + rloc) * a.ldc + col0 + lane, v);
          }
        }
      }
      prefetchOld(col0 + 32);
      prefetchGate(col0 + 32);
    }
    __syncwarp();  // staging buffer is reused by the next 32-column block
  }
}


// ---- epilogue through TMA tensor stores -------------------------------------------------------------
// The register -> staging -> 128-bit global store epilogue above costs ~1 us per 32-column block and warp
// (measured with the stamps of the persistent kernel: 4 us per 128 x 128 tile against a 2 us main loop).
// Here a lane keeps ITS ROW of the block (tcgen05.ld 32x32b: lane = row, registers = 32 columns), applies
// alpha / bias (/ swish'(H)), writes the row into a 4 KB shared-memory tile in the 128-byte-swizzled layout
// the C tensor map describes (16-byte chunk j of row r at j ^ (r % 8): conflict free), and ONE thread hands
// the tile to the copy engine: cp.async.bulk.tensor store, or cp.reduce.async.bulk.tensor .add for C += tile
// (accumulating products and split-K partial sums - no read of the old C by the SM at all).  The copy engine
// clips rows / columns beyond the tensor, so ragged tiles need no special path.  Two tiles per warp alternate.
__device__ __forceinline__ void tmaStoreTile(const CUtensorMap* map, const void* src, int c0, int c1, int c2, bool reduceAdd) {
  if(reduceAdd)
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)map), "r"(smemAddr(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)map), "r"(smemAddr(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__device__ __forceinline__ void tmemLoad32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
}

// The 32-column blocks of a tile are software pipelined: the TMEM read of block c + 1 is in flight while block c is
// scaled, written to shared memory and handed to the copy engine (fully unrolled: two register sets alternate).
template <int BN, bool GATE>
__device__ __forceinline__ void epilogueTileTma(const TcArgs& a, const CUtensorMap* tmC, uint32_t tmemBase, uint64_t* tmemFullBar, uint8_t* sbuf, int warp, int lane, int m0, int n0, int batch, int split, uint32_t parity) {
  static_assert(!GATE, "the gated epilogue takes the staging path");
  constexpr int NB = BN / 32;
  const int q = warp & 3;
  const int rowBase = m0 + q * 32;
  const bool addBias = a.bias != nullptr && split == 0;
  const bool reduceAdd = a.tmaStore == 2;
  const uint32_t tbase = tmemBase + ((uint32_t)(q * 32) << 16);

  // bias of the tile's columns: lane l keeps column 32 c + l of every block c, requested BEFORE the accumulator wait
  // (the kernels run with the maximal shared-memory carve-out: there is next to no L1, a load issued inside the
  // block loop costs a full L2 round trip per block - measured: 230 us instead of 140 us for the vocabulary projection)
  float breg[NB];
#pragma unroll
  for(int c = 0; c < NB; ++c)
    breg[c] = (addBias && n0 + c * 32 + lane < a.N) ? __ldg(a.bias + n0 + c * 32 + lane) : 0.f;

  mbarWait(tmemFullBar, parity);
  tcgenFenceAfter();
  if(rowBase >= a.M)
    return;
  uint32_t r[2][32];
  tmemLoad32(tbase, r[0]);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for(int c = 0; c < NB; ++c) {
    const int col0 = n0 + c * 32;
    if(col0 < a.N) {  // (warp-uniform)
      if(c + 1 < NB && col0 + 32 < a.N)
        tmemLoad32(tbase + (uint32_t)((c + 1) * 32), r[(c + 1) & 1]);  // lands while this block is processed
      // the tile buffer used two blocks ago must have been read by the copy engine
      uint8_t* tile = sbuf + (c & 1) * 4096;
      if(c >= 2) {
        if(lane == 0)
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
      }
      const uint32_t rowAddr = smemAddr(tile) + (uint32_t)lane * 128u;
#pragma unroll
      for(int j = 0; j < 8; ++j) {
        float4 v;
        v.x = a.alpha * __uint_as_float(r[c & 1][4 * j]);
        v.y = a.alpha * __uint_as_float(r[c & 1][4 * j + 1]);
        v.z = a.alpha * __uint_as_float(r[c & 1][4 * j + 2]);
        v.w = a.alpha * __uint_as_float(r[c & 1][4 * j + 3]);
        if(addBias) {  // (warp-uniform; columns beyond N carry 0 and are clipped by the store anyway)
          v.x += __shfl_sync(0xffffffffu, breg[c], 4 * j);
          v.y += __shfl_sync(0xffffffffu, breg[c], 4 * j + 1);
          v.z += __shfl_sync(0xffffffffu, breg[c], 4 * j + 2);
          v.w += __shfl_sync(0xffffffffu, breg[c], 4 * j + 3);
        }
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowAddr + (uint32_t)((j ^ (lane & 7)) << 4)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the copy engine
      __syncwarp();
      if(lane == 0)
        tmaStoreTile(tmC, tile, col0, rowBase, batch, reduceAdd);
      if(c + 1 < NB && col0 + 32 < a.N)
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
  }
  // shared memory must stay intact until the copy engine has read the last tiles
  if(lane == 0)
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

// Operand descriptors of a launch.  G > 1: K-grouped product C = sum_g A_g op(B_g) over G separate
// tensor pairs (the input gradient of several projections of the same tensor: dX = dQ Wq^T + dK Wk^T +
// dV Wv^T is ONE launch whose CTAs walk 3 x K/32 k-blocks instead of three dependent launches).
template <int G>
struct alignas(64) TfMaps {
  CUtensorMap a[G];
  CUtensorMap b[G];
  CUtensorMap c;  // fp32 output tile map {32 columns, 32 rows}, SWIZZLE_128B (bf16 kernels, TcArgs::tmaStore)
};

template <int BN, int STAGES, bool A_MN, bool B_MN, int G = 1, bool GATE = false>
__global__ void __launch_bounds__(192, GATE ? 2 : 1) gGemmTf32(const __grid_constant__ TfMaps<G> tm, TcArgs a) {
  typedef TfSmem<BN, STAGES> L;
  extern __shared__ uint8_t smemRaw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
  uint64_t* fullBar = (uint64_t*)(smem + L::BAR_OFFSET);
  uint64_t* emptyBar = fullBar + STAGES;
  uint64_t* tmemFullBar = emptyBar + STAGES;
  uint32_t* tmemHolder = (uint32_t*)(tmemFullBar + 1);

  pdlTrigger();  // the next kernel may start launching; it waits for our completion in its own pdlWait()
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long* stamp = a.stamps ? a.stamps + 5 * ((size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z)) : nullptr;
  auto now = [] {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
  };
  if(stamp && threadIdx.x == 0)
    stamp[0] = now();  // CTA start
  if(a.spanMin && threadIdx.x == 0)
    atomicMin(a.spanMin, now());

  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * BN;
  const int batch = blockIdx.z / a.splits;
  const int split = blockIdx.z - batch * a.splits;
  const int kb0 = split * a.kBlocksPerSplit;
  const int nkb = min(a.kBlocksPerSplit, a.kBlocks - kb0);
  // Bias gradients for free: the CTAs of the first tile column also sum the columns of every A tile
  // they stream (A = the adjoint whose column sums are the bias gradient) with the four warps that
  // otherwise idle until the epilogue.
  const bool doSums = !A_MN && a.colSum[0] != nullptr && blockIdx.y == 0;

  if(warp == 0 && lane == 0) {
#pragma unroll
    for(int g = 0; g < G; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.a[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.b[g]) : "memory");
    }
    for(int s = 0; s < STAGES; ++s) {
      mbarInit(fullBar + s, 1);
      mbarInit(emptyBar + s, doSums ? 5 : 1);  // + the four warps that read the A tile for its column sums
    }
    mbarInit(tmemFullBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if(warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(tmemHolder)), "r"((uint32_t)BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgenFenceBefore();
  __syncthreads();
  tcgenFenceAfter();
  const uint32_t tmemBase = *tmemHolder;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail
  // of the previous kernel; operands and C are only touched from here on
  pdlWait();
  if(stamp && threadIdx.x == 0)
    stamp[1] = now();  // prologue done (barriers, TMEM, descriptors)

  if(warp == 0) {
    if(lane == 0) {
      // ---------------- TMA producer ----------------
      const int batchA = a.rowsPerBatchA ? batch : 0;  // rowsPerBatch* != 0 marks a batched operand
      const int batchB = a.rowsPerBatchB ? batch : 0;
      for(int i = 0; i < nkb; ++i) {
        int s = i % STAGES;
        uint32_t phase = (uint32_t)(i / STAGES) & 1u;
        mbarWait(emptyBar + s, phase ^ 1u);
        mbarExpectTx(fullBar + s, (uint32_t)L::STAGE_BYTES);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        const int kAbs = kb0 + i;
        const int grp = G == 1 ? 0 : kAbs / a.kBlocksGroup;
        const int kc = (G == 1 ? kAbs : kAbs - grp * a.kBlocksGroup) * TF_BLOCK_K;
        const CUtensorMap* tmA = &tm.a[grp];
        const CUtensorMap* tmB = &tm.b[grp];
        if(A_MN) {
#pragma unroll
          for(int c = 0; c < BLOCK_M / 32; ++c)
            tmaLoad3D(tmA, fullBar + s, sa + c * 4096, m0 + 32 * c, kc, batchA);
        } else {
          tmaLoad3D(tmA, fullBar + s, sa, kc, m0, batchA);
        }
        if(B_MN) {
#pragma unroll
          for(int c = 0; c < BN / 32; ++c)
            tmaLoad3D(tmB, fullBar + s, sb + c * 4096, n0 + 32 * c, kc, batchB);
        } else {
          tmaLoad3D(tmB, fullBar + s, sb, kc, n0, batchB);
        }
      }
    }
  } else if(warp == 1) {
    if(lane == 0) {
      // ---------------- MMA issuer (single thread) ----------------
      constexpr uint32_t idesc = makeInstrDescTf32(BLOCK_M, BN, A_MN, B_MN);
      constexpr uint32_t stepA = A_MN ? (1024 >> 4) : (32 >> 4);
      constexpr uint32_t stepB = B_MN ? (1024 >> 4) : (32 >> 4);
      for(int i = 0; i < nkb; ++i) {
        int s = i % STAGES;
        uint32_t phase = (uint32_t)(i / STAGES) & 1u;
        mbarWait(fullBar + s, phase);
        if(stamp && i == 0)
          stamp[2] = now();  // first operands landed
        tcgenFenceAfter();
        uint32_t sa = smemAddr(smem + s * L::STAGE_BYTES);
        uint32_t sb = sa + L::A_BYTES;
        uint64_t descA = makeSmemDescTf32<A_MN>(sa);
        uint64_t descB = makeSmemDescTf32<B_MN>(sb);
#pragma unroll
        for(int k = 0; k < TF_BLOCK_K / 8; ++k)
          ummaTf32(tmemBase, descA + (uint64_t)(k * stepA), descB + (uint64_t)(k * stepB), idesc, (uint32_t)((i | k) != 0));
        ummaCommit(emptyBar + s);
      }
      ummaCommit(tmemFullBar);
    }
  } else {
    float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * kStagePitch);
    if(stamp && threadIdx.x == 64 && !doSums) {  // (with column sums the readers must not stall: the stages wait for them)
      mbarWait(tmemFullBar, 0);
      stamp[3] = now();  // accumulator complete
    }
    if(doSums) {
      // K-major SWIZZLE_128B tile: row R at (R / 8) * 1024 + (R % 8) * 128 bytes, its 16-byte chunk c at
      // position c ^ (R % 8).  lane = column of the k-block, warp = 32 rows: conflict-free LDS.32.
      const int rw = (warp - 2) * 32;
      for(int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        mbarWait(fullBar + s, (uint32_t)(i / STAGES) & 1u);
        const float* sa = reinterpret_cast<const float*>(smem + s * L::STAGE_BYTES);
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for(int r = 0; r < 32; r += 2) {
          const int R0 = rw + r, R1 = R0 + 1;
          acc0 += sa[(R0 >> 3) * 256 + (R0 & 7) * 32 + ((((lane >> 2) ^ (R0 & 7)) << 2) | (lane & 3))];
          acc1 += sa[(R1 >> 3) * 256 + (R1 & 7) * 32 + ((((lane >> 2) ^ (R1 & 7)) << 2) | (lane & 3))];
        }
        const int kAbs = kb0 + i;
        const int grp = G == 1 ? 0 : kAbs / a.kBlocksGroup;
        const int col = (G == 1 ? kAbs : kAbs - grp * a.kBlocksGroup) * TF_BLOCK_K + lane;
        if(col < a.colSumLen)
          atomicAdd(a.colSum[grp] + col, acc0 + acc1);
        __syncwarp();
        if(lane == 0)
          mbarArrive(emptyBar + s);
      }
      // the epilogue stages its tile over stage 0: every reader must be done with it
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    epilogueTile<BN, GATE>(a, tmemBase, tmemFullBar, stage, warp, lane, m0, n0, batch, split);
  }

  tcgenFenceBefore();
  __syncthreads();
  if(stamp && threadIdx.x == 0)
    stamp[4] = now();  // epilogue done
  if(a.spanMax && threadIdx.x == 0)
    atomicMax(a.spanMax, now());
  if(warp == 1) {
    __syncwarp();
    tcgenFenceAfter();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)BN));
  }
}

CUtensorMap makeTensorMap(GemmHandle h, const __nv_bfloat16* base, uint64_t rows, uint64_t K, uint32_t boxRows) {
  if(!h->encodeTiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    ABORT_IF(!fn || qres != cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from the driver");
    h->encodeTiled = (GemmContext::EncodeTiledFn)fn;
  }
  CUtensorMap map;
  cuuint64_t gdim[2] = {K, rows};
  cuuint64_t gstride[1] = {K * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, boxRows};
  cuuint32_t estride[2] = {1, 1};
  CUresult rc = h->encodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ABORT_IF(rc != CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code", (int)rc);
  return map;
}

template <int BN, int STAGES>
void launchTc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, int batches) {
  typedef TcSmem<BN, STAGES> L;
  static bool configured = false;
  if(!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(gGemmTcgen05<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  dim3 grid((a.M + BLOCK_M - 1) / BLOCK_M, (a.N + BN - 1) / BN, batches * a.splits);
  gGemmTcgen05<BN, STAGES><<<grid, 192, L::TOTAL, cudaStreamOfEngine()>>>(tmA, tmB, a);
  CUDA_LAUNCH_CHECK();
}

struct GemmProblem {
  Tensor C, A, B, bias;
  int rowsA, colsA, rowsB, colsB;  // per batch, as stored
  int batches;
  size_t strideA, strideB;  // 0 when the operand is shared by all batches
  bool transA, transB;
  float beta, alpha;
  std::vector<Tensor> moreA, moreB;  // K-grouped product: further (A_g, B_g) pairs of the same shapes
  Tensor gate;                       // swish'-gated epilogue: pre-activation with the shape of C
  std::vector<Tensor> colSums;       // per (A_g): receives += column sums of A_g (bias gradient), K-major A only
};

void runSimt(const GemmProblem& p) {
  SimtArgs a;
  a.A = p.A->data();
  a.B = p.B->data();
  a.C = p.C->data();
  a.bias = p.bias ? p.bias->data() : nullptr;
  a.M = p.transA ? p.colsA : p.rowsA;
  a.K = p.transA ? p.rowsA : p.colsA;
  a.N = p.transB ? p.rowsB : p.colsB;
  a.lda = p.colsA;
  a.ldb = p.colsB;
  a.ldc = a.N;
  a.transA = p.transA;
  a.transB = p.transB;
  a.alpha = p.alpha;
  a.beta = p.beta;
  a.strideA = p.strideA;
  a.strideB = p.strideB;
  a.strideC = (size_t)a.M * a.N;
  dim3 grid((a.N + 63) / 64, (a.M + 63) / 64, p.batches);
  gGemmSimt<<<grid, 256, 0, cudaStreamOfEngine()>>>(a);
  CUDA_LAUNCH_CHECK();
}

void runTensorCore(GemmHandle h, const GemmProblem& p) {
  ABORT_IF(p.A->memory()->fp32Skipped || p.B->memory()->fp32Skipped, "packed GEMM path reached with a shadow-only operand (only its bf16 copy exists)");
  const bool x3 = h->mode == GemmMode::BF16X3 || h->mode == GemmMode::TF32;  // (BF16S falls back to plain bf16)
  int M = p.transA ? p.colsA : p.rowsA;
  int K = p.transA ? p.rowsA : p.colsA;
  int N = p.transB ? p.rowsB : p.colsB;
  bool batched = p.batches > 1;

  // tile shape: fill the machine if 128-wide tiles cannot
  // The config-B products are LATENCY bound (one wave of CTAs, each a serial chain prologue ->
  // first TMA round trip -> 16 k-blocks -> epilogue; measured in-graph: 128x128 tiles 19.8 us vs
  // 128x64 tiles 16 us for 3200x512x512): prefer more, narrower CTAs until the wide tiles
  // alone fill the machine.
  long tiles128 = (long)((M + BLOCK_M - 1) / BLOCK_M) * ((N + 127) / 128) * p.batches;
  int BN = (tiles128 >= kNumSMs && N > 64) ? 128 : 64;

  // A operand (rows = M): K-major means "not transposed" for A, transposed source when transA
  int batchesA = p.strideA ? p.batches : 1;
  int batchesB = p.strideB ? p.batches : 1;
  Packed pa = packOperand(h, p.A->data(), p.rowsA, p.colsA, batchesA, p.strideA, p.transA, false, x3, batched ? BLOCK_M : 0);
  // B operand (rows = N): stored [K, N] when !transB -> needs the transposing pack
  Packed pb = packOperand(h, p.B->data(), p.rowsB, p.colsB, batchesB, p.strideB, !p.transB, true, x3, batched ? BN : 0);
  ABORT_IF(pa.Ktotal != pb.Ktotal, "packed operands disagree on K");

  CUtensorMap tmA = makeTensorMap(h, pa.data, (uint64_t)pa.rowsPad * batchesA, (uint64_t)pa.Ktotal, BLOCK_M);
  CUtensorMap tmB = makeTensorMap(h, pb.data, (uint64_t)pb.rowsPad * batchesB, (uint64_t)pb.Ktotal, (uint32_t)BN);

  TcArgs a = {};
  a.C = p.C->data();
  a.bias = p.bias ? p.bias->data() : nullptr;
  a.M = M;
  a.N = N;
  a.ldc = N;
  a.kBlocks = pa.Ktotal / BLOCK_K;
  a.rowsPerBatchA = (batched && p.strideA) ? pa.rowsPad : 0;
  a.rowsPerBatchB = (batched && p.strideB) ? pb.rowsPad : 0;
  a.strideC = (size_t)M * N;
  a.alpha = p.alpha;
  a.beta = p.beta;
  a.stamps = nullptr;
  a.spanMin = a.spanMax = nullptr;

  // split-K when the tile grid cannot fill the SMs but K is long (weight gradients)
  long tiles = (long)((M + BLOCK_M - 1) / BLOCK_M) * ((N + BN - 1) / BN) * p.batches;
  int splits = 1;
  if(!batched && tiles * 2 <= kNumSMs && a.kBlocks >= 8) {
    splits = (int)std::min<long>((kNumSMs * 2 + tiles - 1) / tiles, a.kBlocks / 4);
    splits = std::max(1, std::min(splits, 32));
  }
  a.kBlocksPerSplit = (a.kBlocks + splits - 1) / splits;
  splits = (a.kBlocks + a.kBlocksPerSplit - 1) / a.kBlocksPerSplit;
  a.splits = splits;
  a.atomicOut = splits > 1;
  if(a.atomicOut && p.beta != 1.f) {
    // atomics accumulate onto C: bring C to beta * C first
    using namespace functional;
    if(p.beta == 0.f)
      p.C->set(0);
    else
      Element(_1 = p.beta * _1, p.C);
  }

  ProfileScope prof(2.0 * M * N * K * p.batches);  // algorithmic flops (not the 3x of the split mode)
  if(BN == 128)
    launchTc<128, 3>(tmA, tmB, a, p.batches);
  else
    launchTc<64, 4>(tmA, tmB, a, p.batches);
  prof.finish();
}


// ---- tf32 path: tensor maps over the fp32 tensors themselves ----
// dims (innermost first): {inner, outer, batch}; box {32, boxOuter, 1}
CUtensorMap makeTensorMapF32(GemmHandle h, const float* base, uint64_t inner, uint64_t outer, uint64_t batches, uint64_t pitchElems, uint64_t batchStrideElems, uint32_t boxOuter, bool mnMajor) {
  if(!h->encodeTiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    ABORT_IF(!fn || qres != cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from the driver");
    h->encodeTiled = (GemmContext::EncodeTiledFn)fn;
  }
  CUtensorMap map;
  cuuint64_t gdim[3] = {inner, outer, batches};
  cuuint64_t gstride[2] = {pitchElems * sizeof(float), (batches > 1 ? batchStrideElems : pitchElems * outer) * sizeof(float)};
  cuuint32_t box[3] = {(cuuint32_t)TF_BLOCK_K, boxOuter, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  // TFLOAT32: the copy engine rounds fp32 -> tf32 (nearest) on the way into shared memory
  CUresult rc = h->encodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, (void*)base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               mnMajor ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ABORT_IF(rc != CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32) failed with code", (int)rc);
  return map;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int G, bool GATE = false>
void launchTf32Maps(const TfMaps<G>& tm, const TcArgs& a, int batches) {
  typedef TfSmem<BN, STAGES> L;
  static bool configured = false;
  if(!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(gGemmTf32<BN, STAGES, A_MN, B_MN, G, GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  dim3 grid((a.M + BLOCK_M - 1) / BLOCK_M, (a.N + BN - 1) / BN, batches * a.splits);
  launchPdl(gGemmTf32<BN, STAGES, A_MN, B_MN, G, GATE>, grid, dim3(192), (size_t)L::TOTAL, cudaStreamOfEngine(), tm, a);
}

// swish'-gated epilogue (both operands K-major)
inline void launchTf32Gated(int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a) {
  TfMaps<1> tm;
  tm.a[0] = tmA;
  tm.b[0] = tmB;
  if(BN == 128)
    launchTf32Maps<128, 3, false, false, 1, true>(tm, a, 1);
  else
    launchTf32Maps<64, 4, false, false, 1, true>(tm, a, 1);
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
void launchTf32(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, int batches) {
  TfMaps<1> tm;
  tm.a[0] = tmA;
  tm.b[0] = tmB;
  launchTf32Maps<BN, STAGES, A_MN, B_MN, 1>(tm, a, batches);
}

// K-grouped launch (both operands K-major: C = sum_g A_g B_g^T)
template <int G>
void launchTf32Grouped(int BN, const TfMaps<G>& tm, const TcArgs& a) {
  if(BN == 128)
    launchTf32Maps<128, 3, false, false, G>(tm, a, 1);
  else
    launchTf32Maps<64, 4, false, false, G>(tm, a, 1);
}

template <bool A_MN, bool B_MN>
void launchTf32Tile(int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, int batches) {
  if(BN == 128)
    launchTf32<128, 3, A_MN, B_MN>(tmA, tmB, a, batches);
  else
    launchTf32<64, 4, A_MN, B_MN>(tmA, tmB, a, batches);
}

inline bool tmaUsable(const float* p, int cols, size_t batchStride) {
  return (((uintptr_t)p) & 15) == 0 && (cols & 3) == 0 && (batchStride & 3) == 0;
}

// Returns false when an operand cannot be described by a tensor map (row pitch or
// address not 16-byte aligned); the caller then takes the packed path.
bool runTf32(GemmHandle h, const GemmProblem& p) {
  if(!tmaUsable(p.A->data(), p.colsA, p.strideA) || !tmaUsable(p.B->data(), p.colsB, p.strideB))
    return false;
  int M = p.transA ? p.colsA : p.rowsA;
  int K = p.transA ? p.rowsA : p.colsA;
  int N = p.transB ? p.rowsB : p.colsB;
  bool batched = p.batches > 1;
  const bool aMN = p.transA;   // stored [K, M]: M contiguous
  const bool bMN = !p.transB;  // stored [K, N]: N contiguous
  const int G = 1 + (int)p.moreA.size();
  if(G > 1) {
    // K-grouped: both operands K-major, whole k-blocks per group, identical shapes, at most 3 pairs
    if(aMN || bMN || batched || G > 3 || (K % TF_BLOCK_K) != 0 || p.moreB.size() != p.moreA.size())
      return false;
    for(int g = 0; g + 1 < G; ++g)
      if(p.moreA[g]->shape() != p.A->shape() || p.moreB[g]->shape() != p.B->shape() || !tmaUsable(p.moreA[g]->data(), p.colsA, 0) || !tmaUsable(p.moreB[g]->data(), p.colsB, 0))
        return false;
  }
  if(p.gate && (aMN || bMN || batched || G > 1))
    return false;
  const int kGroup = (K + TF_BLOCK_K - 1) / TF_BLOCK_K;  // k-blocks of one (A, B) pair

  // Tile width and split-K are picked together by a small cost model calibrated on this GPU
  // (scripts/gemm_probe.py, in-graph timings in profiles/): a CTA costs a fixed prologue, its
  // k-blocks (operand traffic L2 -> smem: 24 KB per block at BN=64, 32 KB at BN=128), an epilogue
  // proportional to the tile; CTAs run 2 per SM in waves of 296; splitting K adds the red.add
  // traffic and, when C is not being accumulated into, a memset.
  const int kBlocksAll = G * kGroup;
  int BN = 64, splits = 1;
  {
    const long mTiles = (M + BLOCK_M - 1) / BLOCK_M;
    double best = 1e30;
    for(int bn : {64, 128}) {
      if(bn == 128 && N <= 64)
        continue;
      const long tiles = mTiles * ((N + bn - 1) / bn) * p.batches;
      const int maxSplits = (batched || kBlocksAll < 32) ? 1 : std::min(32, kBlocksAll / 8);
      for(int sp = 1; sp <= maxSplits; ++sp) {
        const long ctas = tiles * sp;
        const double waves = (double)((ctas + 2 * kNumSMs - 1) / (2 * kNumSMs));
        const double kb = (double)((kBlocksAll + sp - 1) / sp);
        double cta = 3.0 + kb * (bn == 128 ? 0.333 : 0.25) + (bn == 128 ? 3.0 : 1.5);
        double cost = waves * cta;
        if(sp > 1)
          cost += 3.0 + 0.3 * sp + (p.beta == 1.f ? 0.0 : 3.0);
        if(cost < best) {
          best = cost;
          BN = bn;
          splits = sp;
        }
      }
    }
  }
  if(const char* forced = std::getenv("MRN_GEMM_BN"))  // tuning aid (scripts/gemm_probe.py)
    BN = std::atoi(forced) == 128 ? 128 : 64;
  // tuning aid: "MRN_GEMM_TRY=M,N,K,beta,BN,splits" overrides the choice for one problem shape
  if(const char* t = std::getenv("MRN_GEMM_TRY")) {
    int m, n, k, bn, sp;
    float be;
    if(sscanf(t, "%d,%d,%d,%f,%d,%d", &m, &n, &k, &be, &bn, &sp) == 6 && m == M && n == N && k == K * G && be == p.beta && !batched) {
      BN = bn == 128 ? 128 : 64;
      splits = std::max(1, sp);
    }
  }

  uint64_t batchesA = p.strideA ? p.batches : 1, batchesB = p.strideB ? p.batches : 1;
  // stored matrices are [rows, cols] row-major: inner = cols, outer = rows
  CUtensorMap tmA = makeTensorMapF32(h, p.A->data(), (uint64_t)p.colsA, (uint64_t)p.rowsA, batchesA, (uint64_t)p.colsA, p.strideA, aMN ? TF_BLOCK_K : BLOCK_M, aMN);
  CUtensorMap tmB = makeTensorMapF32(h, p.B->data(), (uint64_t)p.colsB, (uint64_t)p.rowsB, batchesB, (uint64_t)p.colsB, p.strideB, bMN ? TF_BLOCK_K : (uint32_t)BN, bMN);

  TcArgs a = {};
  a.C = p.C->data();
  a.bias = p.bias ? p.bias->data() : nullptr;
  a.M = M;
  a.N = N;
  a.ldc = N;
  a.kBlocks = kBlocksAll;
  a.kBlocksGroup = kGroup;
  a.gate = p.gate ? p.gate->data() : nullptr;
  a.colSum[0] = a.colSum[1] = a.colSum[2] = nullptr;
  a.colSumLen = K;
  if(!p.colSums.empty()) {
    ABORT_IF(aMN || batched || (int)p.colSums.size() != G, "column sums need a K-major, unbatched A operand per group");
    for(int g = 0; g < G; ++g) {
      ABORT_IF((int)p.colSums[g]->size() != K, "column-sum target has the wrong length");
      a.colSum[g] = p.colSums[g]->data();
    }
  }
  a.stamps = g_stampBuffer;  // null unless gemmDebugStamps() armed it
  a.rowsPerBatchA = (batched && p.strideA) ? 1 : 0;  // batched-operand flags for the producer
  a.rowsPerBatchB = (batched && p.strideB) ? 1 : 0;
  a.strideC = (size_t)M * N;
  a.alpha = p.alpha;
  a.beta = p.beta;

  long tiles = (long)((M + BLOCK_M - 1) / BLOCK_M) * ((N + BN - 1) / BN) * p.batches;
  (void)tiles;
  if(const char* forced = std::getenv("MRN_GEMM_SPLITS"))
    splits = std::max(1, std::min(std::atoi(forced), a.kBlocks));
  a.kBlocksPerSplit = (a.kBlocks + splits - 1) / splits;
  splits = (a.kBlocks + a.kBlocksPerSplit - 1) / a.kBlocksPerSplit;
  a.splits = splits;
  a.atomicOut = splits > 1;
  if(a.atomicOut && p.beta != 1.f) {
    using namespace functional;
    if(p.beta == 0.f)
      p.C->set(0);
    else
      Element(_1 = p.beta * _1, p.C);
  }

  ProfileScope prof(2.0 * M * N * K * G * p.batches);
  a.spanMin = prof.spanMin;
  a.spanMax = prof.spanMax;
  if(p.gate) {
    launchTf32Gated(BN, tmA, tmB, a);
  } else if(G > 1) {
    TfMaps<3> tm3;
    tm3.a[0] = tmA;
    tm3.b[0] = tmB;
    for(int g = 1; g < G; ++g) {
      tm3.a[g] = makeTensorMapF32(h, p.moreA[g - 1]->data(), (uint64_t)p.colsA, (uint64_t)p.rowsA, 1, (uint64_t)p.colsA, 0, BLOCK_M, false);
      tm3.b[g] = makeTensorMapF32(h, p.moreB[g - 1]->data(), (uint64_t)p.colsB, (uint64_t)p.rowsB, 1, (uint64_t)p.colsB, 0, (uint32_t)BN, false);
    }
    if(G == 2) {
      TfMaps<2> tm2;
      for(int g = 0; g < 2; ++g) {
        tm2.a[g] = tm3.a[g];
        tm2.b[g] = tm3.b[g];
      }
      launchTf32Grouped<2>(BN, tm2, a);
    } else {
      launchTf32Grouped<3>(BN, tm3, a);
    }
  } else if(aMN && bMN)
    launchTf32Tile<true, true>(BN, tmA, tmB, a, p.batches);
  else if(aMN)
    launchTf32Tile<true, false>(BN, tmA, tmB, a, p.batches);
  else if(bMN)
    launchTf32Tile<false, true>(BN, tmA, tmB, a, p.batches);
  else
    launchTf32Tile<false, false>(BN, tmA, tmB, a, p.batches);
  if(prof.on)
    prof.finish(std::to_string(M) + "," + std::to_string(N) + "," + std::to_string(K * G) + "," + std::to_string(p.batches) + "," + (aMN ? "T" : "N") + (bMN ? "N" : "T") + ","
                + std::to_string(BN) + "," + std::to_string(splits) + "," + std::to_string(p.beta));
  return true;
}


// =============================================================================
// bf16 kernel on shadow operands (GemmMode::BF16S)
// =============================================================================
// Same anatomy as gGemmTf32 (TMA producer warp, single-thread MMA issuer, four epilogue warps that
// double as column-sum readers, K-grouped and swish'-gated variants), on 2-byte operands:
//   k-block  = 64 reduction elements = 128 bytes per operand row -> half the bytes per MAC of the
//              tf32 path, kind::f16 MMAs (K = 16 per instruction) at twice the tf32 rate;
//   K-major  operand (reduction dim contiguous): ONE TMA box {64 k, ROWS}, SWIZZLE_128B rows of 128 B,
//              descriptor SBO = 1024 B (8 rows), k-step (16 bf16 = 32 B) advances the start address;
//   MN-major operand (M / N dim contiguous: weights in the forward product, both operands of
//              dW = X^T dY): ROWS/64 TMA boxes {64 mn, 64 k}; each box is 64 k-rows of 128 bytes,
//              SWIZZLE_128B, i.e. 8 canonical MN-major atoms (64 mn x 8 k) stacked along k;
//              descriptor LBO = 8192 B (next 64 mn = next box), SBO = 1024 B (next 8 k-rows),
//              k-step (16 k-rows) = 2048 B.
constexpr int BF_BLOCK_K = 64;

template <bool MN>
__device__ __forceinline__ uint64_t makeSmemDescBf16(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  if(MN)
    d |= (uint64_t)(8192 >> 4) << 16;  // LBO: next group of 64 mn
  else
    d |= (uint64_t)1 << 16;            // unused for K-major SWIZZLE_128B
  d |= (uint64_t)(1024 >> 4) << 32;    // SBO: next 8 rows (K-major) / next 8 k-rows (MN-major)
  d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t makeInstrDescBf16(int M, int N, bool aMN, bool bMN) {
  return (1u << 4)     // c_format = F32
         | (1u << 7)   // a_format = BF16
         | (1u << 10)  // b_format = BF16
         | ((aMN ? 1u : 0u) << 15) | ((bMN ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int G = 1, bool GATE = false>
__global__ void __launch_bounds__(192, GATE ? 2 : 1) gGemmBf16(const __grid_constant__ TfMaps<G> tm, TcArgs a) {
  typedef TfSmem<BN, STAGES> L;  // 128 bytes per operand row and stage, as in the tf32 kernel
  extern __shared__ uint8_t smemRaw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
  uint64_t* fullBar = (uint64_t*)(smem + L::BAR_OFFSET);
  uint64_t* emptyBar = fullBar + STAGES;
  uint64_t* tmemFullBar = emptyBar + STAGES;
  uint32_t* tmemHolder = (uint32_t*)(tmemFullBar + 1);

  pdlTrigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto now = [] {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
  };
  if(a.spanMin && threadIdx.x == 0)
    atomicMin(a.spanMin, now());

  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * BN;
  const int batch = blockIdx.z / a.splits;
  const int split = blockIdx.z - batch * a.splits;
  const int kb0 = split * a.kBlocksPerSplit;
  const int nkb = min(a.kBlocksPerSplit, a.kBlocks - kb0);
  // column sums of the A operand (bias gradients): the A tiles of a tile row pass through every CTA of that row, so
  // the k-blocks are dealt out over the tile columns (k-block kb belongs to column kb mod gridDim.y) instead of
  // leaving all of them to column 0 - those CTAs were the slowest of the grid (0.65 instead of 0.28 us per k-block)
  const bool doSums = !A_MN && a.colSum[0] != nullptr;

  if(warp == 0 && lane == 0) {
#pragma unroll
    for(int g = 0; g < G; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.a[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.b[g]) : "memory");
    }
    for(int s = 0; s < STAGES; ++s) {
      mbarInit(fullBar + s, 1);
      mbarInit(emptyBar + s, doSums ? 5 : 1);
    }
    mbarInit(tmemFullBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if(warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(tmemHolder)), "r"((uint32_t)BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgenFenceBefore();
  __syncthreads();
  tcgenFenceAfter();
  const uint32_t tmemBase = *tmemHolder;
  pdlWait();

  if(warp == 0) {
    if(lane == 0) {
      // ---------------- TMA producer ----------------
      const int batchA = a.rowsPerBatchA ? batch : 0;
      const int batchB = a.rowsPerBatchB ? batch : 0;
      for(int i = 0; i < nkb; ++i) {
        int s = i % STAGES;
        uint32_t phase = (uint32_t)(i / STAGES) & 1u;
        mbarWait(emptyBar + s, phase ^ 1u);
        mbarExpectTx(fullBar + s, (uint32_t)L::STAGE_BYTES);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        const int kAbs = kb0 + i;
        const int grp = G == 1 ? 0 : kAbs / a.kBlocksGroup;
        const int kc = (G == 1 ? kAbs : kAbs - grp * a.kBlocksGroup) * BF_BLOCK_K;
        const CUtensorMap* tmA = &tm.a[grp];
        const CUtensorMap* tmB = &tm.b[grp];
        if(A_MN) {
#pragma unroll
          for(int c = 0; c < BLOCK_M / 64; ++c)
            tmaLoad3D(tmA, fullBar + s, sa + c * 8192, m0 + 64 * c, kc, batchA);
        } else {
          tmaLoad3D(tmA, fullBar + s, sa, kc, m0, batchA);
        }
        if(B_MN) {
#pragma unroll
          for(int c = 0; c < BN / 64; ++c)
            tmaLoad3D(tmB, fullBar + s, sb + c * 8192, n0 + 64 * c, kc, batchB);
        } else {
          tmaLoad3D(tmB, fullBar + s, sb, kc, n0, batchB);
        }
      }
    }
  } else if(warp == 1) {
    if(lane == 0) {
      // ---------------- MMA issuer (single thread) ----------------
      constexpr uint32_t idesc = makeInstrDescBf16(BLOCK_M, BN, A_MN, B_MN);
      constexpr uint32_t stepA = A_MN ? (2048 >> 4) : (32 >> 4);
      constexpr uint32_t stepB = B_MN ? (2048 >> 4) : (32 >> 4);
      for(int i = 0; i < nkb; ++i) {
        int s = i % STAGES;
        uint32_t phase = (uint32_t)(i / STAGES) & 1u;
        mbarWait(fullBar + s, phase);
        tcgenFenceAfter();
        uint32_t sa = smemAddr(smem + s * L::STAGE_BYTES);
        uint32_t sb = sa + L::A_BYTES;
        uint64_t descA = makeSmemDescBf16<A_MN>(sa);
        uint64_t descB = makeSmemDescBf16<B_MN>(sb);
#pragma unroll
        for(int k = 0; k < BF_BLOCK_K / UMMA_K; ++k)
          umma(tmemBase, descA + (uint64_t)(k * stepA), descB + (uint64_t)(k * stepB), idesc, (uint32_t)((i | k) != 0));
        ummaCommit(emptyBar + s);
      }
      ummaCommit(tmemFullBar);
    }
  } else {
    float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * kStagePitch);
    if(doSums) {
      // K-major SWIZZLE_128B tile of bf16: row R at (R / 8) * 1024 + (R % 8) * 128 bytes, its 16-byte chunk c
      // at position c ^ (R % 8).  lane = two neighbouring columns (one 32-bit word), warp = 32 rows.
      const int rw = (warp - 2) * 32;
      for(int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        mbarWait(fullBar + s, (uint32_t)(i / STAGES) & 1u);
        const int kAbs = kb0 + i;
        if(kAbs % (int)gridDim.y == (int)blockIdx.y) {
          const uint32_t* sa = reinterpret_cast<const uint32_t*>(smem + s * L::STAGE_BYTES);
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for(int r = 0; r < 32; ++r) {
            const int R = rw + r;
            const uint32_t w = sa[(R >> 3) * 256 + (R & 7) * 32 + ((((lane >> 2) ^ (R & 7)) << 2) | (lane & 3))];
            acc0 += __uint_as_float(w << 16);
            acc1 += __uint_as_float(w & 0xffff0000u);
          }
          const int grp = G == 1 ? 0 : kAbs / a.kBlocksGroup;
          const int col = (G == 1 ? kAbs : kAbs - grp * a.kBlocksGroup) * BF_BLOCK_K + 2 * lane;
          if(col < a.colSumLen)
            atomicAdd(a.colSum[grp] + col, acc0);
          if(col + 1 < a.colSumLen)
            atomicAdd(a.colSum[grp] + col + 1, acc1);
        }
        __syncwarp();
        if(lane == 0)
          mbarArrive(emptyBar + s);  // (every epilogue warp acknowledges every stage: emptyBar counts 5)
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    if(!GATE && a.tmaStore)  // (the tile buffers overlay the first ring stages: dead once the accumulator is complete)
      epilogueTileTma<BN, false>(a, &tm.c, tmemBase, tmemFullBar, smem + (warp - 2) * 8192, warp, lane, m0, n0, batch, split);
    else
      epilogueTile<BN, GATE>(a, tmemBase, tmemFullBar, stage, warp, lane, m0, n0, batch, split);
  }

  tcgenFenceBefore();
  __syncthreads();
  if(a.spanMax && threadIdx.x == 0)
    atomicMax(a.spanMax, now());
  if(warp == 1) {
    __syncwarp();
    tcgenFenceAfter();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)BN));
  }
}


// ---- persistent variant: one CTA per SM walks a static list of 128 x BN output tiles ----------------
// For products with many tiles (feed-forward layers: 400, vocabulary projection: 6250 / 1000) the
// one-tile-per-CTA kernel pays its serial chain (prologue, first TMA round trip, epilogue) per tile
// and runs in ragged waves.  Here the three roles loop over the tiles independently:
//   producer   streams k-blocks of tile after tile through ONE smem ring (no drain between tiles);
//   MMA issuer accumulates tile j into TMEM buffer j & 1 (2 x BN columns allocated) - it only waits
//              for the epilogue of tile j - 2;
//   epilogue   warps drain buffer j & 1 (own staging memory, the ring stays live) while the main
//              loop of tile j + 1 runs, then hand the buffer back (tmemEmpty).
// Tiles are numbered m-fastest, so the CTAs running at the same time share B tiles through L2.
template <int BN, int STAGES>
struct BfPersistSmem {
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;
  // per epilogue warp: two 4 KB tiles for the TMA stores (>= the 32 x kStagePitch floats of the plain epilogue);
  // five-stage rings leave room for the eight staging areas of the EPI = 8 variant
  static constexpr int STAGING_BYTES = STAGES >= 6 ? 4 * 8192 : 8 * 32 * kStagePitch * 4 + 4096;
  static constexpr int BAR_OFFSET = STAGING_OFFSET + STAGING_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;
};

// SUMS: two more warps follow the ring and add the column sums of the (K-major) A tiles into a.colSum[0] - the bias
// gradient, as in the one-tile kernels (there the idle epilogue warps do it; here they are busy with the previous
// tile); the k-blocks of a tile row are dealt out over its tile columns.  They acknowledge every stage (emptyBar counts 3).
// EPI = 8: two epilogue warps per TMEM lane quarter, each takes half of the tile's columns (the gated epilogue
// evaluates swish'(H) for every element: with four warps it is 4x the main loop).
template <int BN, int STAGES, bool A_MN, bool B_MN, bool GATE = false, bool SUMS = false, int EPI = 4>
__global__ void __launch_bounds__(64 + 32 * EPI + (SUMS ? 64 : 0), 1) gGemmBf16Persistent(const __grid_constant__ TfMaps<1> tm, TcArgs a) {
  typedef BfPersistSmem<BN, STAGES> L;
  static_assert(EPI == 4 || EPI == 8, "four or eight epilogue warps");
  static_assert(EPI * 32 * kStagePitch * 4 <= L::STAGING_BYTES, "staging area too small");
  extern __shared__ uint8_t smemRaw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
  uint64_t* fullBar = (uint64_t*)(smem + L::BAR_OFFSET);
  uint64_t* emptyBar = fullBar + STAGES;
  uint64_t* tmemFullBar = emptyBar + STAGES;  // [2]
  uint64_t* tmemEmptyBar = tmemFullBar + 2;   // [2]
  uint32_t* tmemHolder = (uint32_t*)(tmemEmptyBar + 2);

  pdlTrigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto now = [] {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
  };
  if(a.spanMin && threadIdx.x == 0)
    atomicMin(a.spanMin, now());

  const int mTiles = (a.M + BLOCK_M - 1) / BLOCK_M;
  const int nTiles = (a.N + BN - 1) / BN;
  const int numTiles = mTiles * nTiles;
  const int nkb = a.kBlocks;
  // tuning aid (scripts/gemm_stamps_bf16.py): 64 slots per CTA - [0] start, [1] prologue done, then per tile j < 10:
  // [2+6j] first TMA issued, [+1] first operands seen by the MMA thread, [+2] last MMA issued, [+3] accumulator
  // complete (epilogue woke up), [+4] epilogue done, [+5] last TMA issued; [63] end
  unsigned long long* stamp = a.stamps ? a.stamps + 64 * (size_t)blockIdx.x : nullptr;
  if(stamp && threadIdx.x == 0)
    stamp[0] = now();

  if(warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.b[0]) : "memory");
    for(int s = 0; s < STAGES; ++s) {
      mbarInit(fullBar + s, 1);
      mbarInit(emptyBar + s, SUMS ? 3 : 1);
    }
    for(int b = 0; b < 2; ++b) {
      mbarInit(tmemFullBar + b, 1);
      mbarInit(tmemEmptyBar + b, EPI);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if(warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(tmemHolder)), "r"((uint32_t)(2 * BN)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgenFenceBefore();
  __syncthreads();
  tcgenFenceAfter();
  const uint32_t tmemBase = *tmemHolder;
  pdlWait();
  if(stamp && threadIdx.x == 0)
    stamp[1] = now();

  if(warp == 0) {
    if(lane == 0) {
      // ---------------- TMA producer: one stream of k-blocks over all tiles of this CTA ----------------
      uint32_t it = 0, jt = 0;
      for(int tile = blockIdx.x; tile < numTiles; tile += gridDim.x, ++jt) {
        const int m0 = (tile % mTiles) * BLOCK_M;
        const int n0 = (tile / mTiles) * BN;
        for(int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1u;
          mbarWait(emptyBar + s, phase ^ 1u);
          if(stamp && jt < 10 && (i == 0 || i == nkb - 1))
            stamp[2 + 6 * jt + (i == 0 ? 0 : 5)] = now();
          mbarExpectTx(fullBar + s, (uint32_t)L::STAGE_BYTES);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          const int kc = i * BF_BLOCK_K;
          if(A_MN) {
#pragma unroll
            for(int c = 0; c < BLOCK_M / 64; ++c)
              tmaLoad3D(&tm.a[0], fullBar + s, sa + c * 8192, m0 + 64 * c, kc, 0);
          } else {
            tmaLoad3D(&tm.a[0], fullBar + s, sa, kc, m0, 0);
          }
          if(B_MN) {
#pragma unroll
            for(int c = 0; c < BN / 64; ++c)
              tmaLoad3D(&tm.b[0], fullBar + s, sb + c * 8192, n0 + 64 * c, kc, 0);
          } else {
            tmaLoad3D(&tm.b[0], fullBar + s, sb, kc, n0, 0);
          }
        }
      }
    }
  } else if(warp == 1) {
    if(lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = makeInstrDescBf16(BLOCK_M, BN, A_MN, B_MN);
      constexpr uint32_t stepA = A_MN ? (2048 >> 4) : (32 >> 4);
      constexpr uint32_t stepB = B_MN ? (2048 >> 4) : (32 >> 4);
      uint32_t it = 0, j = 0;
      for(int tile = blockIdx.x; tile < numTiles; tile += gridDim.x, ++j) {
        const uint32_t buf = j & 1u, use = j >> 1;
        mbarWait(tmemEmptyBar + buf, (use & 1u) ^ 1u);  // the epilogue of tile j - 2 has drained this buffer
        tcgenFenceAfter();
        const uint32_t tmemD = tmemBase + buf * BN;
        for(int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1u;
          mbarWait(fullBar + s, phase);
          if(stamp && j < 10 && i == 0)
            stamp[2 + 6 * j + 1] = now();
          tcgenFenceAfter();
          const uint32_t sa = smemAddr(smem + s * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
          const uint64_t descA = makeSmemDescBf16<A_MN>(sa);
          const uint64_t descB = makeSmemDescBf16<B_MN>(sb);
#pragma unroll
          for(int k = 0; k < BF_BLOCK_K / UMMA_K; ++k)
            umma(tmemD, descA + (uint64_t)(k * stepA), descB + (uint64_t)(k * stepB), idesc, (uint32_t)((i | k) != 0));
          ummaCommit(emptyBar + s);
        }
        ummaCommit(tmemFullBar + buf);  // accumulator of this tile complete
        if(stamp && j < 10)
          stamp[2 + 6 * j + 2] = now();
      }
    }
  } else if(SUMS && warp >= 2 + EPI) {
    // ---------------- column-sum warps (rows 64 (warp - first) .. + 63 of every A tile of the first tile column) ----------------
    static_assert(!SUMS || !A_MN, "column sums read K-major A tiles");
    const int rw = (warp - (2 + EPI)) * 64;
    uint32_t it = 0;
    for(int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
      // k-block i of a tile row is summed by the tile of column i mod nTiles (all columns stream the same A tiles)
      const int tileCol = tile / mTiles;
      const bool sums = a.colSum[0] != nullptr;
      for(int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        mbarWait(fullBar + s, (it / STAGES) & 1u);
        if(sums && i % nTiles == tileCol) {
          // K-major SWIZZLE_128B bf16 tile: row R at (R / 8) * 1024 + (R % 8) * 128 bytes, 16-byte chunk c at c ^ (R % 8)
          const uint32_t* sa = reinterpret_cast<const uint32_t*>(smem + s * L::STAGE_BYTES);
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 16
          for(int r = 0; r < 64; ++r) {
            const int R = rw + r;
            const uint32_t w = sa[(R >> 3) * 256 + (R & 7) * 32 + ((((lane >> 2) ^ (R & 7)) << 2) | (lane & 3))];
            acc0 += __uint_as_float(w << 16);
            acc1 += __uint_as_float(w & 0xffff0000u);
          }
          const int col = i * BF_BLOCK_K + 2 * lane;
          if(col < a.colSumLen)
            atomicAdd(a.colSum[0] + col, acc0);
          if(col + 1 < a.colSumLen)
            atomicAdd(a.colSum[0] + col + 1, acc1);
        }
        __syncwarp();
        if(lane == 0)
          mbarArrive(emptyBar + s);
      }
    }
  } else {
    // ---------------- epilogue warps ----------------
    float* stage = reinterpret_cast<float*>(smem + L::STAGING_OFFSET) + (warp - 2) * (32 * kStagePitch);
    uint32_t j = 0;
    for(int tile = blockIdx.x; tile < numTiles; tile += gridDim.x, ++j) {
      const int m0 = (tile % mTiles) * BLOCK_M;
      const int n0 = (tile / mTiles) * BN;
      const uint32_t buf = j & 1u, use = j >> 1;
      if(stamp && j < 10 && threadIdx.x == 64) {  // (costs this warp one extra wait; tuning runs only)
        mbarWait(tmemFullBar + buf, use & 1u);
        stamp[2 + 6 * j + 3] = now();
      }
      if(!GATE && EPI == 4 && a.tmaStore)
        epilogueTileTma<BN, false>(a, &tm.c, tmemBase + buf * BN, tmemFullBar + buf, smem + L::STAGING_OFFSET + (warp - 2) * 8192, warp, lane, m0, n0, 0, 0, use & 1u);
      else if(EPI == 8)  // warps 2..5 take the first half of the columns, warps 6..9 the second
        epilogueTile<BN, GATE>(a, tmemBase + buf * BN, tmemFullBar + buf, stage, warp, lane, m0, n0, 0, 0, use & 1u, warp < 6 ? 0 : BN / 2, warp < 6 ? BN / 2 : BN);
      else
        epilogueTile<BN, GATE>(a, tmemBase + buf * BN, tmemFullBar + buf, stage, warp, lane, m0, n0, 0, 0, use & 1u);
      tcgenFenceBefore();
      __syncwarp();
      if(lane == 0)
        mbarArrive(tmemEmptyBar + buf);
      if(stamp && j < 10 && threadIdx.x == 64)
        stamp[2 + 6 * j + 4] = now();
    }
  }

  tcgenFenceBefore();
  __syncthreads();
  if(a.spanMax && threadIdx.x == 0)
    atomicMax(a.spanMax, now());
  if(stamp && threadIdx.x == 0)
    stamp[63] = now();
  if(warp == 1) {
    __syncwarp();
    tcgenFenceAfter();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)(2 * BN)));
  }
}

template <bool A_MN, bool B_MN, bool GATE, bool SUMS = false>
void launchBf16Persistent(const TfMaps<1>& tm, const TcArgs& a) {
  constexpr int BN = 128;
  constexpr int EPI = GATE ? 8 : 4;      // the gated epilogue is compute heavy: eight warps
  constexpr int STAGES = GATE ? 5 : 6;
  typedef BfPersistSmem<BN, STAGES> L;
  static bool configured = false;
  if(!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(gGemmBf16Persistent<BN, STAGES, A_MN, B_MN, GATE, SUMS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = ((a.M + BLOCK_M - 1) / BLOCK_M) * ((a.N + BN - 1) / BN);
  dim3 grid(std::min(tiles, kNumSMs));
  launchPdl(gGemmBf16Persistent<BN, STAGES, A_MN, B_MN, GATE, SUMS, EPI>, grid, dim3(64 + 32 * EPI + (SUMS ? 64 : 0)), (size_t)L::TOTAL, cudaStreamOfEngine(), tm, a);
}

// dims (innermost first): {inner, outer, batch}; box {64, boxOuter, 1}
CUtensorMap makeTensorMapBf16(GemmHandle h, const __nv_bfloat16* base, uint64_t inner, uint64_t outer, uint64_t batches, uint64_t pitchElems, uint64_t batchStrideElems, uint32_t boxOuter) {
  if(!h->encodeTiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    ABORT_IF(!fn || qres != cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from the driver");
    h->encodeTiled = (GemmContext::EncodeTiledFn)fn;
  }
  CUtensorMap map;
  cuuint64_t gdim[3] = {inner, outer, batches};
  cuuint64_t gstride[2] = {pitchElems * sizeof(__nv_bfloat16), (batches > 1 ? batchStrideElems : pitchElems * outer) * sizeof(__nv_bfloat16)};
  cuuint32_t box[3] = {(cuuint32_t)BF_BLOCK_K, boxOuter, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult rc = h->encodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ABORT_IF(rc != CUDA_SUCCESS, "cuTensorMapEncodeTiled (bf16) failed with code", (int)rc);
  return map;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int G, bool GATE = false>
void launchBf16Maps(const TfMaps<G>& tm, const TcArgs& a, int batches) {
  typedef TfSmem<BN, STAGES> L;
  static bool configured = false;
  if(!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(gGemmBf16<BN, STAGES, A_MN, B_MN, G, GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  dim3 grid((a.M + BLOCK_M - 1) / BLOCK_M, (a.N + BN - 1) / BN, batches * a.splits);
  launchPdl(gGemmBf16<BN, STAGES, A_MN, B_MN, G, GATE>, grid, dim3(192), (size_t)L::TOTAL, cudaStreamOfEngine(), tm, a);
}

template <bool A_MN, bool B_MN, int G, bool GATE = false>
void launchBf16Tile(int BN, const TfMaps<G>& tm, const TcArgs& a, int batches) {
  if(BN == 128) {
    launchBf16Maps<128, 3, A_MN, B_MN, G, GATE>(tm, a, batches);
    return;
  }
  // One tile row and a long K (the state products of a recurrent cell: 64 x 3072 x 1024 per time step, a chain of
  // them): a few dozen CTAs, each a serial walk over 8..16 k-blocks.  An eight-stage ring (192 KB) has every k-block
  // of such a CTA in flight at once - measured on config C: 22.46 ms per step against 22.20 with four stages (the
  // 5 us of such a launch are prologue / first round trip / epilogue, not the ring), so it stays opt-in.
  if constexpr(G == 1 && !GATE && !A_MN) {
    static const bool deepRing = std::getenv("MRN_GEMM_DEEP_RING") != nullptr;
    if(deepRing && a.M <= BLOCK_M && a.kBlocksPerSplit >= 6 && a.colSum[0] == nullptr) {
      launchBf16Maps<64, 8, A_MN, B_MN, G, GATE>(tm, a, batches);
      return;
    }
  }
  launchBf16Maps<64, 4, A_MN, B_MN, G, GATE>(tm, a, batches);
}

// fp32 output [batches, rows, cols] as 32 x 32 tiles, 128-byte swizzle (rows of a tile = 128 bytes)
CUtensorMap makeTensorMapC(GemmHandle h, const float* base, uint64_t cols, uint64_t rows, uint64_t batches) {
  CUtensorMap map;
  cuuint64_t gdim[3] = {cols, rows, batches};
  cuuint64_t gstride[2] = {cols * sizeof(float), cols * rows * sizeof(float)};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult rc = h->encodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ABORT_IF(rc != CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32 output) failed with code", (int)rc);
  return map;
}

// a bf16 operand (row pitch = cols elements) can be described by a tensor map
inline bool tmaUsableBf16(const float* p, int cols, size_t batchStride) {
  return (((uintptr_t)p) & 31) == 0 && (cols & 7) == 0 && (batchStride & 7) == 0;
}


// ---- products that share their A operand, in one launch ("N-grouped") ---------------------------------------------
// The query / key / value projections of an attention block multiply the SAME activations by three weight matrices, and
// their weight gradients multiply the same X^T by three adjoints: three launches of 200 (160) CTAs each, the two extra
// ones on the side stream.  Here the tile columns of ONE grid walk the groups: group g = its own B tensor map, bias and
// C tensor map (outputs stay separate tensors); B is always MN-major (a weight [K, N] or an adjoint [rows, N]), A is
// K-major (forward) or MN-major (weight gradient, split-K with TMA reduce-add).  Epilogue: TMA stores only.
template <int NG>
struct alignas(64) NGroupMaps {
  CUtensorMap a;
  CUtensorMap b[NG];
  CUtensorMap c[NG];
};
struct NGroupArgs {
  const float* bias[3];
  int groupN;         // columns of one group
  int tilesPerGroup;  // ceil(groupN / BN)
};

template <int BN, int STAGES, bool A_MN, int NG>
__global__ void __launch_bounds__(192, 1) gGemmBf16NGroup(const __grid_constant__ NGroupMaps<NG> tm, TcArgs a, NGroupArgs ga) {
  typedef TfSmem<BN, STAGES> L;
  extern __shared__ uint8_t smemRaw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
  uint64_t* fullBar = (uint64_t*)(smem + L::BAR_OFFSET);
  uint64_t* emptyBar = fullBar + STAGES;
  uint64_t* tmemFullBar = emptyBar + STAGES;
  uint32_t* tmemHolder = (uint32_t*)(tmemFullBar + 1);

  pdlTrigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto now = [] {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
  };
  if(a.spanMin && threadIdx.x == 0)
    atomicMin(a.spanMin, now());

  const int m0 = blockIdx.x * BLOCK_M;
  const int grp = blockIdx.y / ga.tilesPerGroup;
  const int n0 = (blockIdx.y - grp * ga.tilesPerGroup) * BN;  // column inside the group
  const int split = blockIdx.z;
  const int kb0 = split * a.kBlocksPerSplit;
  const int nkb = min(a.kBlocksPerSplit, a.kBlocks - kb0);

  if(warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.b[grp]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.c[grp]) : "memory");
    for(int s = 0; s < STAGES; ++s) {
      mbarInit(fullBar + s, 1);
      mbarInit(emptyBar + s, 1);
    }
    mbarInit(tmemFullBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if(warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(tmemHolder)), "r"((uint32_t)BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgenFenceBefore();
  __syncthreads();
  tcgenFenceAfter();
  const uint32_t tmemBase = *tmemHolder;
  pdlWait();

  if(warp == 0) {
    if(lane == 0) {
      for(int i = 0; i < nkb; ++i) {
        int s = i % STAGES;
        uint32_t phase = (uint32_t)(i / STAGES) & 1u;
        mbarWait(emptyBar + s, phase ^ 1u);
        mbarExpectTx(fullBar + s, (uint32_t)L::STAGE_BYTES);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        const int kc = (kb0 + i) * BF_BLOCK_K;
        if(A_MN) {
#pragma unroll
          for(int c = 0; c < BLOCK_M / 64; ++c)
            tmaLoad3D(&tm.a, fullBar + s, sa + c * 8192, m0 + 64 * c, kc, 0);
        } else {
          tmaLoad3D(&tm.a, fullBar + s, sa, kc, m0, 0);
        }
#pragma unroll
        for(int c = 0; c < BN / 64; ++c)
          tmaLoad3D(&tm.b[grp], fullBar + s, sb + c * 8192, n0 + 64 * c, kc, 0);
      }
    }
  } else if(warp == 1) {
    if(lane == 0) {
      constexpr uint32_t idesc = makeInstrDescBf16(BLOCK_M, BN, A_MN, true);
      constexpr uint32_t stepA = A_MN ? (2048 >> 4) : (32 >> 4);
      constexpr uint32_t stepB = 2048 >> 4;
      for(int i = 0; i < nkb; ++i) {
        int s = i % STAGES;
        uint32_t phase = (uint32_t)(i / STAGES) & 1u;
        mbarWait(fullBar + s, phase);
        tcgenFenceAfter();
        uint32_t sa = smemAddr(smem + s * L::STAGE_BYTES);
        uint32_t sb = sa + L::A_BYTES;
        uint64_t descA = makeSmemDescBf16<A_MN>(sa);
        uint64_t descB = makeSmemDescBf16<true>(sb);
#pragma unroll
        for(int k = 0; k < BF_BLOCK_K / UMMA_K; ++k)
          umma(tmemBase, descA + (uint64_t)(k * stepA), descB + (uint64_t)(k * stepB), idesc, (uint32_t)((i | k) != 0));
        ummaCommit(emptyBar + s);
      }
      ummaCommit(tmemFullBar);
    }
  } else {
    TcArgs al = a;  // this group's view: its bias, its column count (the C tensor map clips the rest)
    al.bias = ga.bias[grp];
    al.N = ga.groupN;
    epilogueTileTma<BN, false>(al, &tm.c[grp], tmemBase, tmemFullBar, smem + (warp - 2) * 8192, warp, lane, m0, n0, 0, split);
  }

  tcgenFenceBefore();
  __syncthreads();
  if(a.spanMax && threadIdx.x == 0)
    atomicMax(a.spanMax, now());
  if(warp == 1) {
    __syncwarp();
    tcgenFenceAfter();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)BN));
  }
}

template <int BN, int STAGES, bool A_MN, int NG>
void launchBf16NGroup(const NGroupMaps<NG>& tm, const TcArgs& a, const NGroupArgs& ga) {
  typedef TfSmem<BN, STAGES> L;
  static bool configured = false;
  if(!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(gGemmBf16NGroup<BN, STAGES, A_MN, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  dim3 grid((a.M + BLOCK_M - 1) / BLOCK_M, NG * ga.tilesPerGroup, a.splits);
  launchPdl(gGemmBf16NGroup<BN, STAGES, A_MN, NG>, grid, dim3(192), (size_t)L::TOTAL, cudaStreamOfEngine(), tm, a, ga);
}

// Column sums of a bf16 operand [rows, cols] (row-major), added into fp32 sums[cols]: the bias gradient that belongs to
// an input-gradient product dX = adj W^T (column sums of adj).  Taking these sums from the A tiles inside the product
// (the epilogue warps acknowledge every ring stage) costs the non-persistent kernel ~0.3 us per k-block - 10 us of the
// 24 us of the 3200 x 512 x 2048 product (profiles/gemm_colsums_r02.md) - so for those products the sums are a pass of
// their own over the bf16 copy, on the side stream: 13 MB, off the critical path.  blockDim (32, 8): a lane owns 8
// adjacent columns (one 16-byte load per row), the 8 warps stride the rows of a slice, one atomic per column and block.
__global__ void __launch_bounds__(256) gColumnSumsBf16(float* __restrict__ sums, const __nv_bfloat16* __restrict__ in, int rows, int cols, int rowsPerSlice) {
  __shared__ float red[8][32][9];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int r0 = blockIdx.y * rowsPerSlice;
  const int r1 = min(rows, r0 + rowsPerSlice);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if(c < cols) {
    int row = r0 + threadIdx.y;
    for(; row + 8 < r1; row += 16) {  // two rows in flight
      const uint4 q0 = *reinterpret_cast<const uint4*>(in + (size_t)row * cols + c);
      const uint4 q1 = *reinterpret_cast<const uint4*>(in + (size_t)(row + 8) * cols + c);
      const uint32_t w0[4] = {q0.x, q0.y, q0.z, q0.w}, w1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for(int e = 0; e < 4; ++e) {
        acc[2 * e] += __uint_as_float(w0[e] << 16) + __uint_as_float(w1[e] << 16);
        acc[2 * e + 1] += __uint_as_float(w0[e] & 0xffff0000u) + __uint_as_float(w1[e] & 0xffff0000u);
      }
    }
    for(; row < r1; row += 8) {
      const uint4 q0 = *reinterpret_cast<const uint4*>(in + (size_t)row * cols + c);
      const uint32_t w0[4] = {q0.x, q0.y, q0.z, q0.w};
#pragma unroll
      for(int e = 0; e < 4; ++e) {
        acc[2 * e] += __uint_as_float(w0[e] << 16);
        acc[2 * e + 1] += __uint_as_float(w0[e] & 0xffff0000u);
      }
    }
  }
#pragma unroll
  for(int e = 0; e < 8; ++e)
    red[threadIdx.y][threadIdx.x][e] = acc[e];
  __syncthreads();
  // 256 columns of the strip, one per thread
  const int t = threadIdx.y * 32 + threadIdx.x;
  const int col = blockIdx.x * 256 + t;
  if(col < cols) {
    float sum = 0.f;
#pragma unroll
    for(int y = 0; y < 8; ++y)
      sum += red[y][t >> 3][t & 7];
    atomicAdd(sums + col, sum);
  }
}

// on the side stream (a bias gradient: nothing reads it before the optimizer); joins are the caller's (graph) business
void columnSumsBf16OnSide(float* sums, const __nv_bfloat16* in, int rows, int cols) {
  const bool side = device::onSide();
  if(!side)
    device::forkSide();
  int strips = (cols + 255) / 256;
  int slices = std::max(1, std::min((kNumSMs * 2 + strips - 1) / strips, (rows + 31) / 32));
  int rowsPerSlice = (rows + slices - 1) / slices;
  slices = (rows + rowsPerSlice - 1) / rowsPerSlice;
  gColumnSumsBf16<<<dim3(strips, slices), dim3(32, 8), 0, cudaStreamOfEngine()>>>(sums, in, rows, cols, rowsPerSlice);
  CUDA_LAUNCH_CHECK();
  if(!side)
    device::returnFromSide();
}

}  // namespace
void ProdFlushColumnSums(GemmHandle h) {
  if(!h || h->pendingSums.empty())
    return;
  device::setDevice(h->device);
  for(auto& ps : h->pendingSums)
    columnSumsBf16OnSide(ps.sums, ps.in, ps.rows, ps.cols);
  h->pendingSums.clear();
}
namespace {
// Returns false when an operand cannot be described by a tensor map; the caller then takes the packed path.
bool runBf16(GemmHandle h, const GemmProblem& p) {
  if(!tmaUsableBf16(p.A->rawData(), p.colsA, p.strideA) || !tmaUsableBf16(p.B->rawData(), p.colsB, p.strideB))
    return false;
  int M = p.transA ? p.colsA : p.rowsA;
  int K = p.transA ? p.rowsA : p.colsA;
  int N = p.transB ? p.rowsB : p.colsB;
  bool batched = p.batches > 1;
  const bool aMN = p.transA;   // stored [K, M]: M contiguous
  const bool bMN = !p.transB;  // stored [K, N]: N contiguous
  const int G = 1 + (int)p.moreA.size();
  if(G > 1) {
    if(aMN || bMN || batched || G > 3 || (K % BF_BLOCK_K) != 0 || p.moreB.size() != p.moreA.size())
      return false;
    for(int g = 0; g + 1 < G; ++g)
      if(p.moreA[g]->shape() != p.A->shape() || p.moreB[g]->shape() != p.B->shape() || !tmaUsableBf16(p.moreA[g]->rawData(), p.colsA, 0) || !tmaUsableBf16(p.moreB[g]->rawData(), p.colsB, 0))
        return false;
  }
  if(p.gate && (aMN || bMN || batched || G > 1))
    return false;
  const int kGroup = (K + BF_BLOCK_K - 1) / BF_BLOCK_K;
  const int kBlocksAll = G * kGroup;

  // tile width / split-K: the cost model of the tf32 path (a k-block moves the same bytes here)
  int BN = 64, splits = 1;
  {
    const long mTiles = (M + BLOCK_M - 1) / BLOCK_M;
    double best = 1e30;
    for(int bn : {64, 128}) {
      if(bn == 128 && N <= 64)
        continue;
      const long tiles = mTiles * ((N + bn - 1) / bn) * p.batches;
      const int maxSplits = (batched || kBlocksAll < 16) ? 1 : std::min(32, kBlocksAll / 4);
      for(int sp = 1; sp <= maxSplits; ++sp) {
        const long ctas = tiles * sp;
        const double waves = (double)((ctas + 2 * kNumSMs - 1) / (2 * kNumSMs));
        const double kb = (double)((kBlocksAll + sp - 1) / sp);
        double cta = 3.0 + kb * (bn == 128 ? 0.333 : 0.25) + (bn == 128 ? 3.0 : 1.5);
        double cost = waves * cta;
        if(sp > 1)
          cost += 3.0 + 0.3 * sp + (p.beta == 1.f ? 0.0 : 3.0);
        if(cost < best) {
          best = cost;
          BN = bn;
          splits = sp;
        }
      }
    }
  }
  if(const char* forced = std::getenv("MRN_GEMM_BN"))
    BN = std::atoi(forced) == 128 ? 128 : 64;
  if(const char* t = std::getenv("MRN_GEMM_TRY")) {
    int m, n, k, bn, sp;
    float be;
    if(sscanf(t, "%d,%d,%d,%f,%d,%d", &m, &n, &k, &be, &bn, &sp) == 6 && m == M && n == N && k == K * G && be == p.beta && !batched) {
      BN = bn == 128 ? 128 : 64;
      splits = std::max(1, sp);
    }
  }
  if(const char* forced = std::getenv("MRN_GEMM_SPLITS"))
    splits = std::max(1, std::min(std::atoi(forced), kBlocksAll));
  // many tiles (feed-forward, vocabulary projection): the persistent kernel - one CTA per SM, tile loop with a
  // double-buffered accumulator
  static const bool noPersistent = std::getenv("MRN_GEMM_NO_PERSISTENT") != nullptr;
  static const int persistMinTiles = std::getenv("MRN_GEMM_PERSIST_TILES") ? std::atoi(std::getenv("MRN_GEMM_PERSIST_TILES")) : 2 * kNumSMs;
  const long tiles128 = (long)((M + BLOCK_M - 1) / BLOCK_M) * ((N + 127) / 128);
  const bool persistent = !noPersistent && !batched && G == 1 && (p.colSums.empty() || (!aMN && !bMN)) && N >= 128 && tiles128 >= persistMinTiles && (splits == 1 || tiles128 >= 4 * kNumSMs);
  if(persistent) {
    BN = 128;
    splits = 1;
  }

  // bf16 operands: shadows written by the producers, the parameter-arena copy, or converted now
  const __nv_bfloat16* a16 = ensureShadow(h, p.A);
  const __nv_bfloat16* b16 = ensureShadow(h, p.B);
  uint64_t batchesA = p.strideA ? p.batches : 1, batchesB = p.strideB ? p.batches : 1;
  TfMaps<3> tm3;
  tm3.a[0] = makeTensorMapBf16(h, a16, (uint64_t)p.colsA, (uint64_t)p.rowsA, batchesA, (uint64_t)p.colsA, p.strideA, aMN ? BF_BLOCK_K : BLOCK_M);
  tm3.b[0] = makeTensorMapBf16(h, b16, (uint64_t)p.colsB, (uint64_t)p.rowsB, batchesB, (uint64_t)p.colsB, p.strideB, bMN ? BF_BLOCK_K : (uint32_t)BN);
  for(int g = 1; g < G; ++g) {
    tm3.a[g] = makeTensorMapBf16(h, ensureShadow(h, p.moreA[g - 1]), (uint64_t)p.colsA, (uint64_t)p.rowsA, 1, (uint64_t)p.colsA, 0, BLOCK_M);
    tm3.b[g] = makeTensorMapBf16(h, ensureShadow(h, p.moreB[g - 1]), (uint64_t)p.colsB, (uint64_t)p.rowsB, 1, (uint64_t)p.colsB, 0, (uint32_t)BN);
  }

  TcArgs a = {};
  a.C = p.C->rawData();  // (replaced below once split-K / shadow-only output are decided)
  a.bias = p.bias ? p.bias->data() : nullptr;
  a.M = M;
  a.N = N;
  a.ldc = N;
  a.kBlocks = kBlocksAll;
  a.kBlocksGroup = kGroup;
  a.gate = p.gate ? p.gate->data() : nullptr;
  a.colSumLen = K;
  if(!p.colSums.empty()) {
    ABORT_IF(aMN || batched || (int)p.colSums.size() != G, "column sums need a K-major, unbatched A operand per group");
    // 0: always inside the product (the A tiles stream through anyway); 1 (default): inside the persistent kernel (its two
    // column-sum warps), a pass of their own for the one-tile kernels (see gColumnSumsBf16); 2: always a pass of their own
    static const int policy = std::getenv("MRN_COLSUM_POLICY") ? std::atoi(std::getenv("MRN_COLSUM_POLICY")) : 1;
    const bool outside = policy == 2 || (policy == 1 && !persistent);
    for(int g = 0; g < G; ++g) {
      ABORT_IF((int)p.colSums[g]->size() != K, "column-sum target has the wrong length");
      if(outside)
        h->pendingSums.push_back({p.colSums[g]->data(), g == 0 ? a16 : ensureShadow(h, p.moreA[g - 1]), M, K});
      else
        a.colSum[g] = p.colSums[g]->data();
    }
  }
  a.rowsPerBatchA = (batched && p.strideA) ? 1 : 0;
  a.rowsPerBatchB = (batched && p.strideB) ? 1 : 0;
  a.strideC = (size_t)M * N;
  a.alpha = p.alpha;
  a.beta = p.beta;
  a.stamps = persistent ? g_stampBuffer : nullptr;  // null unless gemmDebugStamps() armed it
  a.kBlocksPerSplit = (a.kBlocks + splits - 1) / splits;
  splits = (a.kBlocks + a.kBlocksPerSplit - 1) / a.kBlocksPerSplit;
  a.splits = splits;
  a.atomicOut = splits > 1;
  if(a.atomicOut && p.beta != 1.f) {
    using namespace functional;
    if(p.beta == 0.f)
      p.C->set(0);
    else
      Element(_1 = p.beta * _1, p.C);
  }
  // C itself a later product operand (and written here in one piece): leave its bf16 copy as well - and ONLY the
  // bf16 copy when nobody reads the fp32 tensor (shadow-only adjoints: the swish'-gated dH of the feed-forward block)
  a.shadowC = a.atomicOut ? nullptr : shadow::produce(p.C);
  a.C = (a.shadowC && p.beta == 0.f) ? shadow::fp32Target(p.C, a.shadowC) : p.C->data();
  // epilogue through the copy engine: plain store, or reduce-add for C += tile (beta = 1, split-K partial sums)
  static const bool noTmaStore = std::getenv("MRN_GEMM_NO_TMA_STORE") != nullptr;
  if(!noTmaStore && !a.shadowC && !p.gate && a.C && (N & 3) == 0 && (((uintptr_t)a.C) & 15) == 0 && (p.beta == 0.f || p.beta == 1.f || a.atomicOut)) {
    a.tmaStore = (a.atomicOut || p.beta == 1.f) ? 2 : 1;
    tm3.c = makeTensorMapC(h, a.C, (uint64_t)N, (uint64_t)M, (uint64_t)p.batches);
  }

  ProfileScope prof(2.0 * M * N * K * G * p.batches);
  a.spanMin = prof.spanMin;
  a.spanMax = prof.spanMax;
  if(persistent) {
    TfMaps<1> tm;
    tm.a[0] = tm3.a[0];
    tm.b[0] = tm3.b[0];
    tm.c = tm3.c;
    if(p.gate && a.colSum[0])
      launchBf16Persistent<false, false, true, true>(tm, a);
    else if(p.gate)
      launchBf16Persistent<false, false, true>(tm, a);
    else if(a.colSum[0])
      launchBf16Persistent<false, false, false, true>(tm, a);
    else if(aMN && bMN)
      launchBf16Persistent<true, true, false>(tm, a);
    else if(aMN)
      launchBf16Persistent<true, false, false>(tm, a);
    else if(bMN)
      launchBf16Persistent<false, true, false>(tm, a);
    else
      launchBf16Persistent<false, false, false>(tm, a);
  } else if(p.gate) {
    TfMaps<1> tm;
    tm.a[0] = tm3.a[0];
    tm.b[0] = tm3.b[0];
    tm.c = tm3.c;
    launchBf16Tile<false, false, 1, true>(BN, tm, a, 1);
  } else if(G == 3) {
    launchBf16Tile<false, false, 3>(BN, tm3, a, 1);
  } else if(G == 2) {
    TfMaps<2> tm2;
    for(int g = 0; g < 2; ++g) {
      tm2.a[g] = tm3.a[g];
      tm2.b[g] = tm3.b[g];
    }
    tm2.c = tm3.c;
    launchBf16Tile<false, false, 2>(BN, tm2, a, 1);
  } else {
    TfMaps<1> tm;
    tm.a[0] = tm3.a[0];
    tm.b[0] = tm3.b[0];
    tm.c = tm3.c;
    if(aMN && bMN)
      launchBf16Tile<true, true, 1>(BN, tm, a, p.batches);
    else if(aMN)
      launchBf16Tile<true, false, 1>(BN, tm, a, p.batches);
    else if(bMN)
      launchBf16Tile<false, true, 1>(BN, tm, a, p.batches);
    else
      launchBf16Tile<false, false, 1>(BN, tm, a, p.batches);
  }
  if(prof.on)
    prof.finish(std::to_string(M) + "," + std::to_string(N) + "," + std::to_string(K * G) + "," + std::to_string(p.batches) + "," + (aMN ? "T" : "N") + (bMN ? "N" : "T") + ","
                + std::to_string(persistent ? -BN : BN) + "," + std::to_string(splits) + "," + std::to_string(p.beta));
  return true;
}

void runGemm(GemmHandle h, const GemmProblem& problem) {
  GemmProblem p = problem;
  device::setDevice(p.C->getDevice());
  // first writer assigns: an adjoint that is still "lazily zero" is overwritten (beta = 0)
  // instead of being memset and then read back by the epilogue
  if(p.C->takeLazyZero())
    p.beta = 0.f;
  int M = p.transA ? p.colsA : p.rowsA;
  int K = p.transA ? p.rowsA : p.colsA;
  int Kb = p.transB ? p.colsB : p.rowsB;
  int N = p.transB ? p.rowsB : p.colsB;
  ABORT_IF(K != Kb, "matrix product requires dimensions to match", K, Kb);
  ABORT_IF((long)p.C->size() != (long)M * N * p.batches, "Prod: output tensor has the wrong size");
  if(M == 0 || N == 0)
    return;
  if(h->mode == GemmMode::FP32)
    runSimt(p);
  else if(h->mode == GemmMode::TF32) {
    if(!runTf32(h, p))
      runTensorCore(h, p);  // operand not TMA-addressable: packed hi/lo bf16 path (at least as precise)
  } else if(h->mode == GemmMode::BF16S) {
    if(!runBf16(h, p))
      runTensorCore(h, p);  // operand not TMA-addressable: packed bf16 path (same arithmetic)
  } else
    runTensorCore(h, p);
}

// TMA-direct tensor-core modes: the fused variants (K-grouped, gated, column sums) exist for both
inline bool directMode(GemmHandle h) {
  return h->mode == GemmMode::TF32 || h->mode == GemmMode::BF16S;
}
inline bool directUsable(GemmHandle h, const Tensor& t, int cols) {
  return h->mode == GemmMode::BF16S ? tmaUsableBf16(t->rawData(), cols, 0) : tmaUsable(t->rawData(), cols, 0);
}
inline bool runDirect(GemmHandle h, const GemmProblem& p) {
  return h->mode == GemmMode::BF16S ? runBf16(h, p) : runTf32(h, p);
}

}  // namespace

// C = beta C + sum_g A_g B_g^T in one launch (K-grouped tcgen05 product) when the tf32 path can
// take it, otherwise as the chain of accumulating products it replaces.
bool ProdColumnSumsFusable(GemmHandle h, const Tensor A) {
  static const bool enabled = std::getenv("MRN_NO_FUSE_BIAS_GRAD") == nullptr;
  return enabled && directMode(h) && directUsable(h, A, A->shape().back());
}

namespace {
// column sums the tensor-core path did not deliver (fallback products): the plain reduction
void addColumnSums(const std::vector<Tensor>& As, const std::vector<Tensor>& colSums) {
  using namespace functional;
  for(size_t g = 0; g < colSums.size(); ++g)
    Add(_1, colSums[g], As[g]);
}
}  // namespace

void ProdGroupedNT(GemmHandle h, Tensor C, const std::vector<Tensor>& As, const std::vector<Tensor>& Bs, float beta, const std::vector<Tensor>& colSums) {
  ABORT_IF(As.empty() || As.size() != Bs.size(), "ProdGroupedNT: need matching operand lists");
  ABORT_IF(!colSums.empty() && colSums.size() != As.size(), "ProdGroupedNT: one column-sum target per operand pair");
  if(As.size() <= 3 && directMode(h) && (As.size() > 1 || !colSums.empty())) {
    GemmProblem p;
    p.C = C;
    p.A = As[0];
    p.B = Bs[0];
    p.bias = nullptr;
    p.colsA = p.A->shape().back();
    p.rowsA = p.A->shape().elements() / p.colsA;
    p.colsB = p.B->shape().back();
    p.rowsB = p.B->shape().elements() / p.colsB;
    p.batches = 1;
    p.strideA = p.strideB = 0;
    p.transA = false;
    p.transB = true;
    p.beta = beta;
    p.alpha = 1.f;
    p.moreA.assign(As.begin() + 1, As.end());
    p.moreB.assign(Bs.begin() + 1, Bs.end());
    p.colSums = colSums;
    device::setDevice(C->getDevice());
    ABORT_IF(p.colsA != p.colsB, "matrix product requires dimensions to match", p.colsA, p.colsB);
    ABORT_IF((long)C->size() != (long)p.rowsA * p.rowsB, "ProdGroupedNT: output tensor has the wrong size");
    const bool wasLazy = C->takeLazyZero();
    if(wasLazy)
      p.beta = 0.f;
    if(p.rowsA == 0 || p.rowsB == 0)
      return;
    if(runDirect(h, p))
      return;
    if(wasLazy)
      beta = 0.f;  // the mark is consumed: the first product of the chain assigns
  }
  for(size_t g = 0; g < As.size(); ++g)
    Prod(h, C, As[g], Bs[g], false, true, g == 0 ? beta : 1.f, 1.f);
  addColumnSums(As, colSums);
}

// dH = beta dH + (A B^T) o swish'(H): input gradient of an affine layer whose input is swish(H), written
// straight into the adjoint of H.  Only the tf32 tensor-core path implements it.
bool ProdSwishGradFusable(GemmHandle h, const Tensor C, const Tensor A, const Tensor B, const Tensor H) {
  if(!directMode(h) || std::getenv("MRN_NO_SWISH_FUSION"))
    return false;
  int colsA = A->shape().back(), colsB = B->shape().back();
  return colsA == colsB && C->shape().elements() == H->shape().elements() && directUsable(h, A, colsA) && directUsable(h, B, colsB) && (((uintptr_t)H->data()) & 15) == 0
         && (int)(B->shape().elements() / colsB) % 4 == 0;
}

void ProdSwishGradNT(GemmHandle h, Tensor C, const Tensor A, const Tensor B, const Tensor H, float beta, Tensor colSum) {
  GemmProblem p;
  p.C = C;
  p.A = A;
  p.B = B;
  p.bias = nullptr;
  p.gate = H;
  if(colSum)
    p.colSums = {colSum};
  p.colsA = A->shape().back();
  p.rowsA = A->shape().elements() / p.colsA;
  p.colsB = B->shape().back();
  p.rowsB = B->shape().elements() / p.colsB;
  p.batches = 1;
  p.strideA = p.strideB = 0;
  p.transA = false;
  p.transB = true;
  p.beta = beta;
  p.alpha = 1.f;
  device::setDevice(C->getDevice());
  ABORT_IF((long)C->size() != (long)p.rowsA * p.rowsB || C->size() != H->size(), "ProdSwishGradNT: shapes do not match");
  if(C->takeLazyZero())
    p.beta = 0.f;
  ABORT_IF(!directMode(h) || !runDirect(h, p), "ProdSwishGradNT: not supported for these operands (check ProdSwishGradFusable first)");
}

// C_g = beta C_g + op(A) B_g (+ bias_g) for 2 or 3 products that share A, as ONE launch (bf16 shadow mode, TMA epilogue).
// Returns false (nothing done) when the operands do not fit; the caller then issues the products one by one.
bool ProdSharedA(GemmHandle h, const std::vector<Tensor>& Cs, const Tensor A, const std::vector<Tensor>& Bs, const std::vector<Tensor>& biases, bool transA, float beta) {
  static const bool disabled = std::getenv("MRN_NO_NGROUP") != nullptr;
  const int NG = (int)Cs.size();
  if(disabled || h->mode != GemmMode::BF16S || NG < 2 || NG > 3 || Bs.size() != Cs.size() || (!biases.empty() && biases.size() != Cs.size()) || (beta != 0.f && beta != 1.f))
    return false;
  const int colsA = A->shape().back(), rowsA = (int)(A->shape().elements() / colsA);
  const int M = transA ? colsA : rowsA, K = transA ? rowsA : colsA;
  const int N = Bs[0]->shape().back(), rowsB = (int)(Bs[0]->shape().elements() / N);
  if(rowsB != K || (N & 7) != 0 || (colsA & 7) != 0 || !tmaUsableBf16(A->rawData(), colsA, 0))
    return false;
  for(int g = 0; g < NG; ++g) {
    if(Bs[g]->shape() != Bs[0]->shape() || (long)Cs[g]->size() != (long)M * N || Cs[g]->isLazyZero() || !tmaUsableBf16(Bs[g]->rawData(), N, 0) || (((uintptr_t)Cs[g]->rawData()) & 15) != 0
       || Cs[g]->memory()->shadowWanted)
      return false;
    if(!biases.empty() && (int)biases[g]->size() != N)
      return false;
  }
  device::setDevice(Cs[0]->getDevice());
  const int kBlocks = (K + BF_BLOCK_K - 1) / BF_BLOCK_K;
  const long mTiles = (M + BLOCK_M - 1) / BLOCK_M;
  // one wave of two CTAs per SM where possible: 64-wide tiles unless they overflow it, split-K for short grids (weight gradients)
  int BN = (mTiles * NG * ((N + 63) / 64) > 2 * kNumSMs && N >= 128) ? 128 : 64;
  const long tiles = mTiles * NG * ((N + BN - 1) / BN);
  int splits = 1;
  if(beta == 1.f && kBlocks >= 16 && tiles * 2 <= 2 * kNumSMs)
    splits = (int)std::max<long>(1, std::min<long>((2 * kNumSMs) / tiles, kBlocks / 4));

  NGroupMaps<3> tm3;
  tm3.a = makeTensorMapBf16(h, ensureShadow(h, A), (uint64_t)colsA, (uint64_t)rowsA, 1, (uint64_t)colsA, 0, transA ? BF_BLOCK_K : BLOCK_M);
  NGroupArgs ga = {};
  ga.groupN = N;
  ga.tilesPerGroup = (N + BN - 1) / BN;
  for(int g = 0; g < NG; ++g) {
    tm3.b[g] = makeTensorMapBf16(h, ensureShadow(h, Bs[g]), (uint64_t)N, (uint64_t)K, 1, (uint64_t)N, 0, BF_BLOCK_K);
    tm3.c[g] = makeTensorMapC(h, Cs[g]->data(), (uint64_t)N, (uint64_t)M, 1);
    ga.bias[g] = biases.empty() ? nullptr : biases[g]->data();
  }
  TcArgs a = {};
  a.M = M;
  a.N = N;
  a.ldc = N;
  a.kBlocks = kBlocks;
  a.kBlocksGroup = kBlocks;
  a.kBlocksPerSplit = (kBlocks + splits - 1) / splits;
  splits = (kBlocks + a.kBlocksPerSplit - 1) / a.kBlocksPerSplit;
  a.splits = splits;
  a.atomicOut = splits > 1;
  a.alpha = 1.f;
  a.beta = beta;
  a.tmaStore = (beta == 1.f || splits > 1) ? 2 : 1;
  ABORT_IF(splits > 1 && beta != 1.f, "ProdSharedA: split-K needs an accumulating product");
  ProfileScope prof(2.0 * M * N * K * NG);
  a.spanMin = prof.spanMin;
  a.spanMax = prof.spanMax;
  auto launch = [&](auto ngTag) {
    constexpr int G = decltype(ngTag)::value;
    NGroupMaps<G> tm;
    tm.a = tm3.a;
    for(int g = 0; g < G; ++g) {
      tm.b[g] = tm3.b[g];
      tm.c[g] = tm3.c[g];
    }
    if(BN == 128) {
      if(transA)
        launchBf16NGroup<128, 3, true, G>(tm, a, ga);
      else
        launchBf16NGroup<128, 3, false, G>(tm, a, ga);
    } else {
      if(transA)
        launchBf16NGroup<64, 4, true, G>(tm, a, ga);
      else
        launchBf16NGroup<64, 4, false, G>(tm, a, ga);
    }
  };
  if(NG == 2)
    launch(std::integral_constant<int, 2>());
  else
    launch(std::integral_constant<int, 3>());
  if(prof.on)
    prof.finish(std::to_string(M) + "," + std::to_string(N * NG) + "," + std::to_string(K) + ",1," + (transA ? "T" : "N") + "N," + std::to_string(BN) + "," + std::to_string(splits) + "," + std::to_string(beta));
  return true;
}

void Prod(GemmHandle h, Tensor C, const Tensor A, const Tensor B, bool transA, bool transB, float beta, float scalar) {
  GemmProblem p;
  p.C = C;
  p.A = A;
  p.B = B;
  p.bias = nullptr;
  p.colsA = A->shape().back();
  p.rowsA = A->shape().elements() / p.colsA;
  p.colsB = B->shape().back();
  p.rowsB = B->shape().elements() / p.colsB;
  p.batches = 1;
  p.strideA = p.strideB = 0;
  p.transA = transA;
  p.transB = transB;
  p.beta = beta;
  p.alpha = scalar;
  runGemm(h, p);
}

void ProdAffine(GemmHandle h, Tensor C, const Tensor A, const Tensor B, const Tensor bias) {
  GemmProblem p;
  p.C = C;
  p.A = A;
  p.B = B;
  p.bias = bias;
  p.colsA = A->shape().back();
  p.rowsA = A->shape().elements() / p.colsA;
  p.colsB = B->shape().back();
  p.rowsB = B->shape().elements() / p.colsB;
  ABORT_IF((int)bias->size() != p.colsB, "ProdAffine: bias must be a row vector of the output width");
  p.batches = 1;
  p.strideA = p.strideB = 0;
  p.transA = p.transB = false;
  p.beta = 0.f;
  p.alpha = 1.f;
  runGemm(h, p);
}

void ProdBatched(GemmHandle h, Tensor C, const Tensor A, const Tensor B, bool transA, bool transB, float beta, float scalar) {
  GemmProblem p;
  p.C = C;
  p.A = A;
  p.B = B;
  p.bias = nullptr;
  p.rowsA = A->shape()[-2];
  p.colsA = A->shape()[-1];
  p.rowsB = B->shape()[-2];
  p.colsB = B->shape()[-1];
  size_t batchA = A->shape().elements() / ((size_t)p.rowsA * p.colsA);
  size_t batchB = B->shape().elements() / ((size_t)p.rowsB * p.colsB);
  p.batches = (int)std::max(batchA, batchB);
  p.strideA = batchA == 1 ? 0 : (size_t)p.rowsA * p.colsA;
  p.strideB = batchB == 1 ? 0 : (size_t)p.rowsB * p.colsB;
  if(p.batches == 1)
    p.strideA = p.strideB = 0;
  p.transA = transA;
  p.transB = transB;
  p.beta = beta;
  p.alpha = scalar;
  runGemm(h, p);
}

}  // namespace marian
