// Kernel-to-kernel cost inside a CUDA graph (B200): a chain of N dependent launches of a persistent-style kernel
// (148 CTAs x 384 threads, each CTA spins `work` clocks), replayed from a graph, with and without programmatic
// dependent launch (PDL: cudaLaunchAttributeProgrammaticStreamSerialization + griddepcontrol.wait in the kernel).
// per-launch time - work = what one more launch in the step costs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_gap launch_gap.cu
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return -1.f; } } while (0)

template <int PDL>
__global__ void __launch_bounds__(384, 1) k(long long work, int* sink, int smem_touch) {
  extern __shared__ int sm[];
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next grid start its prologue
  if (smem_touch) sm[threadIdx.x] = threadIdx.x;                             // prologue stand-in
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");                // inputs of this grid are now complete
  const long long t0 = clock64();
  while (clock64() - t0 < work) {
  }
  if (threadIdx.x == 0 && work < 0) *sink = sm[0];
}

static float run(int pdl, int n, long long work, size_t smem, int* sink) {
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && i > 0) ? 1 : 0;     // the first node has no kernel predecessor
    const int touch = 1;
    if (pdl) CK(cudaLaunchKernelEx(&cfg, k<1>, work, sink, touch));
    else CK(cudaLaunchKernelEx(&cfg, k<0>, work, sink, touch));
  }
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  CK(cudaGraphLaunch(ge, s));
  CK(cudaStreamSynchronize(s));
  cudaEventRecord(e0, s);
  for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  CK(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaStreamDestroy(s);
  return ms / 5 / n * 1e3f;   // us per launch
}

int main() {
  int* sink;
  if (cudaMalloc(&sink, 4) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
    printf("cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  const int n = 1000;
  for (size_t smem : {(size_t)2048, (size_t)200 * 1024}) {
    for (long long work : {0LL, 20000LL, 100000LL}) {   // 0, ~10 us, ~50 us of work per kernel
      const float a = run(0, n, work, smem, sink), b = run(1, n, work, smem, sink);
      printf("smem %3zu KiB  work %6lld clk (%5.1f us): plain %6.2f us/launch   PDL %6.2f us/launch\n", smem >> 10, work,
             work / 1965.0, a, b);
      fflush(stdout);
    }
  }
  return 0;
}
