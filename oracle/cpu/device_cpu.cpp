// TEST ORACLE — not part of the product.  Host-memory implementation of the
// engine's device services (csrc/tensors/device.h) so that the host graph code
// can be linked against the CPU restatement of the tensor operators.
#include <cstdlib>
#include <cstring>

#include "common/definitions.h"
#include "tensors/device.h"

namespace marian {
namespace device {

void setDevice(int) {}
int getDevice() { return 0; }
void* currentStream() { return nullptr; }
void setStream(void*) {}

void* mallocDevice(size_t bytes) {
  void* p = nullptr;
  if(posix_memalign(&p, 256, bytes ? bytes : 256) != 0)
    ABORT("oracle: out of host memory", bytes);
  return p;
}
void freeDevice(void* p) { std::free(p); }
void* mallocPinned(size_t bytes) { return mallocDevice(bytes); }
void freePinned(void* p) { std::free(p); }

void* pinnedScratch(size_t bytes) {
  static thread_local void* buf = nullptr;
  static thread_local size_t cap = 0;
  if(bytes > cap) {
    std::free(buf);
    cap = bytes < 4096 ? 4096 : bytes;
    buf = mallocDevice(cap);
  }
  return buf;
}
void copyH2D(void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); }
void copyH2DBlocking(void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); }
void copyD2H(void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); }
void copyD2D(void* dst, const void* src, size_t bytes) { std::memmove(dst, src, bytes); }
void zero(void* dst, size_t bytes) { std::memset(dst, 0, bytes); }
void fill(float* dst, float value, size_t n) {
  for(size_t i = 0; i < n; ++i)
    dst[i] = value;
}
void synchronize() {}
size_t ipcHandleBytes() { return 64; }
void ipcExport(void*, unsigned char*) {}
void* ipcOpen(const unsigned char*) { return nullptr; }
void forkSide() {}
void returnFromSide() {}
void joinSide() {}
bool onSide() { return false; }
// lanes: one host thread runs everything in tape order
void openLanes() {}
void closeLanes() {}
bool lanesOpen() { return false; }
void selectLane(int) {}
int currentLane() { return 0; }
void* laneMark() { return nullptr; }
void laneWait(void*) {}
void* recordMarker(void*) { return nullptr; }
void waitMarker(void*) {}
void freeMarker(void*) {}
bool capturing() { return false; }

bool captureSupported() { return false; }
void beginCapture() { ABORT("oracle: capture not supported"); }
void* endCapture() { return nullptr; }
size_t lastCaptureKernelCount() { return 0; }
void launchGraph(void*) {}
void destroyGraph(void*) {}

const char* backendName() { return "cpu-oracle"; }

}  // namespace device
}  // namespace marian
