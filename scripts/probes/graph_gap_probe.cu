// Micro-benchmark: cost of one node of a dependent chain inside a CUDA graph on this GPU.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/probes/graph_gap_probe.cu -o gpurun_out/graph_gap_probe
// Variants: plain kernel chain; chain with the programmatic-serialization attribute (PDL) and the
// trigger at the start / at the end / absent; chain interleaved with external event-record nodes
// (what bench.py's per-launch GEMM timing inserts); chain interleaved with small memset nodes.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t err__ = (x); if(err__ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(err__), __LINE__); return 1; } } while(0)

template <int MODE>  // 0 plain, 1 trigger early, 2 trigger late, 3 wait only
__global__ void work(float* p, int n, int iters) {
  if(MODE == 1) asm volatile("griddepcontrol.launch_dependents;");
  if(MODE != 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) {
    float v = p[i];
    for(int k = 0; k < iters; ++k) v = v * 1.0001f + 0.5f;
    p[i] = v;
  }
  if(MODE == 2) asm volatile("griddepcontrol.launch_dependents;");
}

template <int MODE>
cudaError_t launch(float* p, int n, int blocks, int iters, cudaStream_t s, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = s;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = a; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, work<MODE>, p, n, iters);
}

int main() {
  const int N = 1 << 20, CHAIN = 400;
  float* d; CK(cudaMalloc(&d, N * 4)); CK(cudaMemset(d, 0, N * 4));
  cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  std::vector<cudaEvent_t> evs(2 * CHAIN);
  for(auto& ev : evs) CK(cudaEventCreate(&ev));
  struct V { const char* name; int mode; bool pdl; int extra; int blocks; int iters; };
  V variants[] = {{"plain, 148 blocks, ~2us of work", 0, false, 0, 148, 2000},      {"PDL trigger early", 1, true, 0, 148, 2000},
                  {"PDL trigger late", 2, true, 0, 148, 2000},                      {"PDL wait only", 3, true, 0, 148, 2000},
                  {"plain + 2 external event nodes per kernel", 0, false, 1, 148, 2000}, {"plain + 1 memset node per kernel", 0, false, 2, 148, 2000},
                  {"plain, tiny kernels (1 block, no work)", 0, false, 0, 1, 1},     {"PDL trigger early, tiny kernels", 1, true, 0, 1, 1}};
  for(auto& v : variants) {
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
    for(int i = 0; i < CHAIN; ++i) {
      if(v.extra == 1) CK(cudaEventRecordWithFlags(evs[2 * i], s, cudaEventRecordExternal));
      if(v.extra == 2) CK(cudaMemsetAsync(d + N - 64, 0, 256, s));
      cudaError_t rc = v.mode == 0 ? launch<0>(d, N, v.blocks, v.iters, s, v.pdl) : v.mode == 1 ? launch<1>(d, N, v.blocks, v.iters, s, v.pdl)
                       : v.mode == 2 ? launch<2>(d, N, v.blocks, v.iters, s, v.pdl) : launch<3>(d, N, v.blocks, v.iters, s, v.pdl);
      CK(rc);
      if(v.extra == 1) CK(cudaEventRecordWithFlags(evs[2 * i + 1], s, cudaEventRecordExternal));
    }
    CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for(int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(a, s));
    const int REP = 20;
    for(int r = 0; r < REP; ++r) CK(cudaGraphLaunch(ge, s));
    CK(cudaEventRecord(b, s)); CK(cudaStreamSynchronize(s));
    float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
    printf("{\"variant\": \"%s\", \"us_per_kernel_node\": %.3f}\n", v.name, ms * 1000.0 / (REP * CHAIN));
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  }
  return 0;
}
