// Device memory + stream services used by the host engine.
//
// The reference hard-wires cudaMalloc/cudaMemcpy + cudaStreamSynchronize(0)
// into TensorBase and DeviceGPU (src/tensors/tensor.cu:21-74,
// src/tensors/device_gpu.cu:17-36).  Here every device interaction of the
// host engine goes through this small interface so that
//   * the product build (device_gpu.cu) issues everything asynchronously on
//     ONE explicit stream, which makes a whole training step capturable into
//     a CUDA graph, and
//   * the test oracle (oracle/cpu/device_cpu.cpp) can link the same host
//     graph code against plain host memory.
#pragma once

#include <cstddef>
#include <cstdint>

namespace marian {
namespace device {

// Selects the CUDA device for the calling thread (no-op on the CPU oracle).
void setDevice(int deviceId);
int getDevice();

// The stream all engine work of the calling thread is issued on.  The product
// build returns a cudaStream_t (as void*).  setStream(nullptr) restores the
// engine-owned stream.
void* currentStream();
void setStream(void* stream);

void* mallocDevice(size_t bytes);
void freeDevice(void* p);
void* mallocPinned(size_t bytes);
void freePinned(void* p);
// Thread-local pinned bounce buffer of at least `bytes` (grows, never shrinks):
// cudaHostAlloc/cudaFreeHost cost ~0.1-100 ms per call depending on the host, so
// blocking read-backs must not allocate.  Valid until the next call on this thread.
void* pinnedScratch(size_t bytes);

// All asynchronous on currentStream().  `src` of copyH2D must be pinned memory
// that stays valid until the copy ran (see tensors/staging.h).
void copyH2D(void* dst, const void* pinnedSrc, size_t bytes);
void copyD2H(void* pinnedDst, const void* src, size_t bytes);
void copyD2D(void* dst, const void* src, size_t bytes);
void zero(void* dst, size_t bytes);
void fill(float* dst, float value, size_t n);

// Blocks the host until currentStream() is drained.
void synchronize();

// True while currentStream() is being captured into a CUDA graph; host code
// uses it to refuse operations that need a host round-trip.
bool capturing();

// Blocking upload from pageable host memory (large one-off initialisations).
void copyH2DBlocking(void* dst, const void* src, size_t bytes);

// ---- whole-step capture (CUDA graphs) -----------------------------------
// captureSupported() is false on the CPU oracle.  beginCapture() starts
// recording everything issued on currentStream(); endCapture() returns an
// executable graph handle (nullptr if recording failed, the work was then NOT
// executed and the caller must run it eagerly).
bool captureSupported();
void beginCapture();
void* endCapture();
// kernel nodes in the graph returned by the last endCapture() of this thread
size_t lastCaptureKernelCount();
void launchGraph(void* exec);
void destroyGraph(void* exec);

// ---- host-visible progress markers ------------------------------------------------------------
// recordMarker(m): records "everything issued so far on currentStream()" into marker m (a fresh one when m is
// null) and returns it; waitMarker(m) blocks the host until that point has been reached.  Used to keep the host
// from refilling pinned staging that uploads of an earlier, still queued step have not read yet.
void* recordMarker(void* marker);
void waitMarker(void* marker);
void freeMarker(void* marker);

// ---- side stream: work off the critical path ---------------------------------
// The backward sweep is a long dependency chain of small kernels; weight and bias gradients
// hang off that chain (nothing reads them before the optimizer).  forkSide() routes the engine
// work this thread issues next to a second stream, ordered after everything issued so far on
// the main stream; returnFromSide() switches back (no ordering); joinSide() makes the main
// stream wait for all side work.  Under capture the side stream joins the recording, so the
// replayed graph has the same two branches.  No-ops on the CPU oracle.
void forkSide();
void returnFromSide();
void joinSide();
// true between forkSide() and returnFromSide()/joinSide()
bool onSide();

// ---- lanes: independent chains of one pass -------------------------------------------------------
// The two directional stacks of a bidirectional recurrent encoder are two long chains of small,
// dependent kernels that share nothing but their input; on one stream they run back to back.  The
// graph tags nodes with a lane (ExpressionGraph::setLane); while a lane region is open, work of lane
// k > 0 goes to that lane's own stream.  openLanes() starts a region (an event on the main stream that
// every lane waits for at its first use); closeLanes() makes the main stream wait for every lane used
// since and ends the region.  laneMark() / laneWait() order individual results across lanes: a mark
// stands for "everything issued so far on the current lane", laneWait makes the current lane wait for
// it (marks of an earlier region are complete by construction and ignored; nullptr is ignored).  The
// side stream forks from the CURRENT lane.  Under capture the lanes join the recording, so the
// replayed CUDA graph has the same parallel branches.  All no-ops on the CPU oracle.
constexpr int kMaxLanes = 4;
void openLanes();
void closeLanes();
bool lanesOpen();
void selectLane(int lane);  // 0 = main stream; ignored (main) outside a region
int currentLane();
void* laneMark();
void laneWait(void* mark);

// ---- inter-process device memory (peer-memory gradient exchange) -----------------------------
size_t ipcHandleBytes();
// handle of the allocation that STARTS at ptr (arenas and the signal pad are whole allocations)
void ipcExport(void* ptr, unsigned char* handleOut);
// maps another process' allocation into this one (enables peer access lazily)
void* ipcOpen(const unsigned char* handle);

// "cuda" for the product, "cpu-oracle" for the test oracle.
const char* backendName();

}  // namespace device
}  // namespace marian
