// CUDA implementation of the engine's device services (tensors/device.h).
//
// One explicit stream per host thread carries ALL engine work: kernels, the
// pinned-staging uploads, cost read-back.  Nothing in here synchronises except
// synchronize() / copyH2DBlocking().  (The reference uses the legacy/per-thread
// default stream and cudaStreamSynchronize(0) after most tensor methods:
// src/tensors/tensor.cu:21-74.)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common/definitions.h"
#include "kernels/cuda_helpers.h"
#include "tensors/device.h"

namespace marian {
namespace device {

namespace {
struct ThreadCtx {
  int device{-1};
  cudaStream_t own{nullptr};   // engine-owned stream of the current device
  cudaStream_t user{nullptr};  // stream injected by the host harness
  bool hasUser{false};
  bool capturing{false};
  size_t lastKernelCount{0};
  // side streams (see device.h): one per lane, so that the off-critical-path work of one chain does not queue
  // behind that of another chain
  cudaStream_t sides[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
  bool sideDirty[kMaxLanes] = {false, false, false, false};
  int sideOf{0};  // lane whose side stream is selected while onSide
  bool onSide{false};
  std::vector<cudaEvent_t> events;  // fork/join markers, reused every step
  size_t nextEvent{0};
  // lanes (see device.h)
  struct LaneMark {
    cudaEvent_t event{nullptr};
    int lane{0};
    uint64_t seq{0};
    uint64_t region{0};
  };
  cudaStream_t lanes[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
  bool laneUsed[kMaxLanes] = {false, false, false, false};
  uint64_t laneSeq[kMaxLanes] = {0, 0, 0, 0};               // marks taken per lane in this region
  uint64_t laneSeen[kMaxLanes][kMaxLanes] = {};             // [waiter][source]: highest mark already waited for
  int lane{0};
  bool lanesOpen{false};
  uint64_t region{0};
  cudaEvent_t laneOpenEvent{nullptr};
  std::vector<LaneMark*> marks;  // pooled per region
  size_t nextMark{0};
};
thread_local ThreadCtx tctx;

cudaStream_t mainStream();

cudaStream_t laneStream() {
  if(tctx.lanesOpen && tctx.lane > 0)
    return tctx.lanes[tctx.lane];
  return mainStream();
}

cudaStream_t stream() {
  if(tctx.onSide)
    return tctx.sides[tctx.sideOf];
  return laneStream();
}

cudaEvent_t nextMarker() {
  if(tctx.nextEvent == tctx.events.size()) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    tctx.events.push_back(e);
  }
  return tctx.events[tctx.nextEvent++];
}

// Stream priorities: the main stream and the lanes carry the dependency chain of the step, the side streams carry
// work nothing waits for before the optimizer (weight / bias gradients, the early exchange phase).  When CTAs of both
// are pending the block scheduler takes the high-priority ones first, so a weight-gradient product that happens to
// run next to an input-gradient product of the chain no longer takes half of the machine from it.  (Kernel nodes of
// a captured graph keep the priority of the stream they were captured on.)  Measured on the Transformer-base step: 4.813
// vs 4.817 ms - no gain, and a low-priority early exchange phase at N = 8 was not measured - so this is OPT-IN:
// MRN_STREAM_PRIORITY=1.
bool usePriorities() {
  static const bool on = std::getenv("MRN_STREAM_PRIORITY") != nullptr;
  return on;
}
cudaStream_t newStream(bool chain) {
  cudaStream_t st = nullptr;
  int least = 0, greatest = 0;
  if(usePriorities() && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess && least != greatest) {
    int pr = greatest;
    if(!chain)
      pr = least;
    else if(tctx.hasUser)  // lanes next to a stream injected by the harness: its priority
      cudaStreamGetPriority(tctx.user, &pr);
    CUDA_CHECK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, pr));
  } else {
    CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  }
  return st;
}

cudaStream_t mainStream() {
  if(tctx.hasUser)
    return tctx.user;
  if(!tctx.own) {
    if(tctx.device < 0)
      setDevice(0);
    tctx.own = newStream(true);
  }
  return tctx.own;
}

__global__ void gFill(float* d, float v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for(; i < n; i += stride)
    d[i] = v;
}
}  // namespace

// One process drives ONE GPU (the multi-GPU groups are one process per GPU); a host thread may still be pointed at
// another device, e.g. by a test harness.  Streams and events belong to the device they were created on: a device
// change drops them (they are re-created lazily), including a stream injected by the harness.  The CUDA runtime's
// current device can also be changed behind our back (torch.cuda.set_device): it is re-asserted when it differs.
void setDevice(int deviceId) {
  if(tctx.device == deviceId) {
    int current = -1;
    if(cudaGetDevice(&current) == cudaSuccess && current != deviceId)
      CUDA_CHECK(cudaSetDevice(deviceId));
    return;
  }
  CUDA_CHECK(cudaSetDevice(deviceId));
  tctx.device = deviceId;
  tctx.own = nullptr;  // streams are (re)created lazily for the new device
  tctx.user = nullptr;
  tctx.hasUser = false;
  tctx.onSide = false;
  for(int k = 0; k < kMaxLanes; ++k) {
    tctx.sides[k] = nullptr;
    tctx.sideDirty[k] = false;
    tctx.lanes[k] = nullptr;
  }
  tctx.lanesOpen = false;
  tctx.lane = 0;
  tctx.laneOpenEvent = nullptr;
  tctx.events.clear();  // events of the previous device are abandoned with its streams
  tctx.nextEvent = 0;
  tctx.marks.clear();
  tctx.nextMark = 0;
}
int getDevice() {
  return tctx.device < 0 ? 0 : tctx.device;
}

void* currentStream() {
  return (void*)stream();
}
void setStream(void* s) {
  tctx.user = (cudaStream_t)s;
  tctx.hasUser = (s != nullptr);
}

void* mallocDevice(size_t bytes) {
  void* p = nullptr;
  CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 256));
  return p;
}
void freeDevice(void* p) {
  if(p)
    CUDA_CHECK(cudaFree(p));
}
void* mallocPinned(size_t bytes) {
  void* p = nullptr;
  CUDA_CHECK(cudaHostAlloc(&p, bytes ? bytes : 256, cudaHostAllocDefault));
  return p;
}
void freePinned(void* p) {
  if(p)
    CUDA_CHECK(cudaFreeHost(p));
}

void* pinnedScratch(size_t bytes) {
  static thread_local void* buf = nullptr;
  static thread_local size_t cap = 0;
  if(bytes > cap) {
    if(buf)
      CUDA_CHECK(cudaFreeHost(buf));
    cap = std::max(bytes, (size_t)1 << 20);
    CUDA_CHECK(cudaHostAlloc(&buf, cap, cudaHostAllocDefault));
  }
  return buf;
}

void copyH2D(void* dst, const void* src, size_t bytes) {
  CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream()));
}
void copyH2DBlocking(void* dst, const void* src, size_t bytes) {
  ABORT_IF(tctx.capturing, "blocking upload during CUDA graph capture");
  CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream()));
  CUDA_CHECK(cudaStreamSynchronize(stream()));
}
void copyD2H(void* dst, const void* src, size_t bytes) {
  CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream()));
}
void copyD2D(void* dst, const void* src, size_t bytes) {
  CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream()));
}
void zero(void* dst, size_t bytes) {
  CUDA_CHECK(cudaMemsetAsync(dst, 0, bytes, stream()));
}
void fill(float* dst, float value, size_t n) {
  if(!n)
    return;
  int threads = 256;
  int blocks = (int)std::min<size_t>((n + threads - 1) / threads, 148 * 8);
  gFill<<<blocks, threads, 0, stream()>>>(dst, value, n);
  CUDA_CHECK(cudaGetLastError());
}

void synchronize() {
  ABORT_IF(tctx.capturing, "stream synchronisation during CUDA graph capture");
  CUDA_CHECK(cudaStreamSynchronize(stream()));
}

bool capturing() {
  return tctx.capturing;
}
bool captureSupported() {
  return true;
}
void beginCapture() {
  ABORT_IF(tctx.capturing, "nested capture");
  CUDA_CHECK(cudaStreamBeginCapture(mainStream(), cudaStreamCaptureModeRelaxed));
  tctx.capturing = true;
}
void* endCapture() {
  ABORT_IF(!tctx.capturing, "endCapture without beginCapture");
  tctx.capturing = false;
  cudaGraph_t graph = nullptr;
  cudaError_t rc = cudaStreamEndCapture(mainStream(), &graph);
  if(rc != cudaSuccess || !graph) {
    fprintf(stderr, "[marian_b200] graph capture failed: %s\n", cudaGetErrorString(rc));
    cudaGetLastError();
    return nullptr;
  }
  {
    size_t n = 0;
    tctx.lastKernelCount = 0;
    if(cudaGraphGetNodes(graph, nullptr, &n) == cudaSuccess && n > 0) {
      std::vector<cudaGraphNode_t> nodes(n);
      cudaGraphGetNodes(graph, nodes.data(), &n);
      for(auto node : nodes) {
        cudaGraphNodeType type;
        if(cudaGraphNodeGetType(node, &type) == cudaSuccess && type == cudaGraphNodeTypeKernel)
          tctx.lastKernelCount++;
        if(std::getenv("MRN_GRAPH_STATS") && cudaGraphNodeGetType(node, &type) == cudaSuccess) {
          static int counts[16];
          if(node == nodes.front())
            for(int& c : counts) c = 0;
          counts[(int)type < 16 ? (int)type : 15]++;
          if(node == nodes.back())
            fprintf(stderr, "[marian_b200] captured graph: %d kernel, %d memcpy, %d memset, %d event-record, %d event-wait, %d other nodes\n", counts[cudaGraphNodeTypeKernel],
                    counts[cudaGraphNodeTypeMemcpy], counts[cudaGraphNodeTypeMemset], counts[cudaGraphNodeTypeEventRecord], counts[cudaGraphNodeTypeWaitEvent],
                    (int)n - counts[cudaGraphNodeTypeKernel] - counts[cudaGraphNodeTypeMemcpy] - counts[cudaGraphNodeTypeMemset] - counts[cudaGraphNodeTypeEventRecord] - counts[cudaGraphNodeTypeWaitEvent]);
        }
      }
    }
  }
  cudaGraphExec_t exec = nullptr;
  rc = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if(rc != cudaSuccess) {
    fprintf(stderr, "[marian_b200] graph instantiate failed: %s\n", cudaGetErrorString(rc));
    cudaGetLastError();
    return nullptr;
  }
  return (void*)exec;
}
size_t lastCaptureKernelCount() {
  return tctx.lastKernelCount;
}
void launchGraph(void* exec) {
  CUDA_CHECK(cudaGraphLaunch((cudaGraphExec_t)exec, stream()));
}
void destroyGraph(void* exec) {
  if(exec)
    cudaGraphExecDestroy((cudaGraphExec_t)exec);
}

void* recordMarker(void* marker) {
  cudaEvent_t e = (cudaEvent_t)marker;
  if(!e)
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(e, mainStream()));
  return e;
}
void waitMarker(void* marker) {
  if(marker)
    CUDA_CHECK(cudaEventSynchronize((cudaEvent_t)marker));
}
void freeMarker(void* marker) {
  if(marker)
    cudaEventDestroy((cudaEvent_t)marker);
}

void forkSide() {
  if(tctx.onSide)
    return;
  const int k = tctx.lanesOpen ? tctx.lane : 0;
  if(!tctx.sides[k])
    tctx.sides[k] = newStream(false);
  cudaEvent_t e = nextMarker();
  CUDA_CHECK(cudaEventRecord(e, laneStream()));  // the side work follows what its lane has issued so far
  CUDA_CHECK(cudaStreamWaitEvent(tctx.sides[k], e, 0));
  tctx.sideOf = k;
  tctx.onSide = true;
  tctx.sideDirty[k] = true;
}
void returnFromSide() {
  tctx.onSide = false;
}
bool onSide() {
  return tctx.onSide;
}
void joinSide() {
  tctx.onSide = false;
  for(int k = 0; k < kMaxLanes; ++k)
    if(tctx.sideDirty[k]) {
      cudaEvent_t e = nextMarker();
      CUDA_CHECK(cudaEventRecord(e, tctx.sides[k]));
      CUDA_CHECK(cudaStreamWaitEvent(mainStream(), e, 0));
      tctx.sideDirty[k] = false;
    }
  tctx.nextEvent = 0;
}

void openLanes() {
  if(tctx.lanesOpen)
    return;
  if(!tctx.laneOpenEvent)
    CUDA_CHECK(cudaEventCreateWithFlags(&tctx.laneOpenEvent, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(tctx.laneOpenEvent, mainStream()));
  tctx.lanesOpen = true;
  tctx.lane = 0;
  ++tctx.region;
  tctx.nextMark = 0;
  for(int i = 0; i < kMaxLanes; ++i) {
    tctx.laneUsed[i] = false;
    tctx.laneSeq[i] = 0;
    for(int j = 0; j < kMaxLanes; ++j)
      tctx.laneSeen[i][j] = 0;
  }
}
void closeLanes() {
  if(!tctx.lanesOpen)
    return;
  ABORT_IF(tctx.onSide, "closeLanes() while on the side stream");
  for(int k = 1; k < kMaxLanes; ++k)
    if(tctx.laneUsed[k]) {
      cudaEvent_t e = nextMarker();
      CUDA_CHECK(cudaEventRecord(e, tctx.lanes[k]));
      CUDA_CHECK(cudaStreamWaitEvent(mainStream(), e, 0));
      tctx.laneUsed[k] = false;
    }
  tctx.lane = 0;
  tctx.lanesOpen = false;
}
bool lanesOpen() {
  return tctx.lanesOpen;
}
void selectLane(int lane) {
  if(!tctx.lanesOpen || lane <= 0 || lane >= kMaxLanes) {
    tctx.lane = 0;
    return;
  }
  if(!tctx.laneUsed[lane]) {
    if(!tctx.lanes[lane])
      tctx.lanes[lane] = newStream(true);
    CUDA_CHECK(cudaStreamWaitEvent(tctx.lanes[lane], tctx.laneOpenEvent, 0));
    tctx.laneUsed[lane] = true;
  }
  tctx.lane = lane;
}
int currentLane() {
  return tctx.lanesOpen ? tctx.lane : 0;
}
void* laneMark() {
  if(!tctx.lanesOpen)
    return nullptr;
  if(tctx.nextMark == tctx.marks.size()) {
    auto* m = new ThreadCtx::LaneMark();
    CUDA_CHECK(cudaEventCreateWithFlags(&m->event, cudaEventDisableTiming));
    tctx.marks.push_back(m);
  }
  auto* m = tctx.marks[tctx.nextMark++];
  m->lane = tctx.lane;
  m->seq = ++tctx.laneSeq[tctx.lane];
  m->region = tctx.region;
  CUDA_CHECK(cudaEventRecord(m->event, laneStream()));
  return m;
}
void laneWait(void* mark) {
  auto* m = (ThreadCtx::LaneMark*)mark;
  if(!m || !tctx.lanesOpen || m->region != tctx.region || m->lane == tctx.lane)
    return;
  if(tctx.laneSeen[tctx.lane][m->lane] >= m->seq)
    return;  // a later mark of that lane was already waited for
  CUDA_CHECK(cudaStreamWaitEvent(laneStream(), m->event, 0));
  tctx.laneSeen[tctx.lane][m->lane] = m->seq;
}

size_t ipcHandleBytes() {
  return sizeof(cudaIpcMemHandle_t);
}
void ipcExport(void* ptr, unsigned char* handleOut) {
  cudaIpcMemHandle_t h;
  CUDA_CHECK(cudaIpcGetMemHandle(&h, ptr));
  std::copy((unsigned char*)&h, (unsigned char*)&h + sizeof(h), handleOut);
}
void* ipcOpen(const unsigned char* handle) {
  cudaIpcMemHandle_t h;
  std::copy(handle, handle + sizeof(h), (unsigned char*)&h);
  void* p = nullptr;
  CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return p;
}

const char* backendName() {
  return "cuda";
}

}  // namespace device
}  // namespace marian
