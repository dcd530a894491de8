// TMEM -> register read bandwidth microbenchmark (sm_100a): how many bytes/clk/SM can tcgen05.ld deliver,
// as a function of resident epilogue-style warps and load shape?   nvcc -arch=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include "../../e4t-diffusion_b200/csrc/common.cuh"
#include <cstdio>
thread_local char g_e4t_err[512];
unsigned long long g_e4t_launches;
int e4t_set_error(const char*, ...) { return 1; }

template <int X>
__global__ void tmem_read_kernel(int iters, int nwarps, unsigned* sink, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      const uint32_t col = (uint32_t)(((i * 32) + (warp >> 2) * 64) & 255);
      if (X == 32) tmem_ld32(base + col, v);
      else { tmem_ld16(base + col, v); tmem_ld16(base + col + 16, v + 16); }
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) acc ^= v[e];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
int main() {
  unsigned* sink; long long* cyc; cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int nw : {4, 8, 16}) {
    for (int x : {32, 16}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (x == 32) tmem_read_kernel<32><<<148, 32 * 16>>>(iters, nw, sink, cyc);
        else tmem_read_kernel<16><<<148, 32 * 16>>>(iters, nw, sink, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)nw * iters * 32 * 32 * 4;
      printf("warps=%2d ld.x%-2d : %lld cycles, %.1f B/clk/SM (%s)\n", nw, x, h, bytes / (double)h, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
