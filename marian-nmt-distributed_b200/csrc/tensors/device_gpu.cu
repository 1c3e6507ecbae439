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
  // side stream (see device.h)
  cudaStream_t side{nullptr};
  bool onSide{false};
  bool sideDirty{false};
  std::vector<cudaEvent_t> events;  // fork/join markers, reused every step
  size_t nextEvent{0};
};
thread_local ThreadCtx tctx;

cudaStream_t mainStream();

cudaStream_t stream() {
  if(tctx.onSide)
    return tctx.side;
  return mainStream();
}

cudaEvent_t nextMarker() {
  if(tctx.nextEvent == tctx.events.size()) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    tctx.events.push_back(e);
  }
  return tctx.events[tctx.nextEvent++];
}

cudaStream_t mainStream() {
  if(tctx.hasUser)
    return tctx.user;
  if(!tctx.own) {
    if(tctx.device < 0)
      setDevice(0);
    CUDA_CHECK(cudaStreamCreateWithFlags(&tctx.own, cudaStreamNonBlocking));
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

void setDevice(int deviceId) {
  if(tctx.device == deviceId)
    return;
  CUDA_CHECK(cudaSetDevice(deviceId));
  tctx.device = deviceId;
  tctx.own = nullptr;  // streams are (re)created lazily for the new device
  tctx.side = nullptr;
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
  cudaStream_t main = mainStream();
  if(!tctx.side)
    CUDA_CHECK(cudaStreamCreateWithFlags(&tctx.side, cudaStreamNonBlocking));
  cudaEvent_t e = nextMarker();
  CUDA_CHECK(cudaEventRecord(e, main));
  CUDA_CHECK(cudaStreamWaitEvent(tctx.side, e, 0));
  tctx.onSide = true;
  tctx.sideDirty = true;
}
void returnFromSide() {
  tctx.onSide = false;
}
bool onSide() {
  return tctx.onSide;
}
void joinSide() {
  tctx.onSide = false;
  if(tctx.sideDirty) {
    cudaEvent_t e = nextMarker();
    CUDA_CHECK(cudaEventRecord(e, tctx.side));
    CUDA_CHECK(cudaStreamWaitEvent(mainStream(), e, 0));
    tctx.sideDirty = false;
  }
  tctx.nextEvent = 0;
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
