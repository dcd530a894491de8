// tcgen05.mma issue-rate microbenchmark (B200): clocks per 128 x N x 16 bf16 MMA as a function of N, for
//   ss      A and B from shared memory, one accumulator (a dependent accumulate chain, as in a GEMM mainloop)
//   ss2     the same, alternating between two accumulators
//   ts      A from tensor memory (as P in P·V), B from shared memory
//   ss-mn   B MN-major (as V in P·V)
//   all SMs / one SM: with 148 CTAs the shared-memory and tensor pipes of every SM run together (power / clocks)
// Question it answers: is there a per-instruction floor that does not shrink with N?  (The GEMM probe shows ~115-125 clk
// per MMA for every tile width below 256, the attention P·V uses N = 48.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../e4t-diffusion_b200/csrc -o mma_rate mma_rate.cu
#include "common.cuh"
#include <stdio.h>

// the two symbols common.cuh expects from c_abi.cu
thread_local char g_e4t_err[512];
unsigned long long g_e4t_launches;
int e4t_set_error(const char*, ...) { return 1; }

#define REPS 512

// mode: 0 ss, 1 ss alternating accumulators, 2 ts, 3 ss with B MN-major; kgroup = MMAs between commits (no waits inside)
__global__ void __launch_bounds__(128, 1) mma_rate(int n, int mode, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  // zero the operands (A 16 KiB, B 32 KiB): denormal / NaN patterns could change the data path's power, not its timing
  for (int i = threadIdx.x; i < (49152 >> 4); i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16((uint32_t)n, false, mode == 3);
    const uint64_t dA = umma_desc(smem_u32(smem), 16, 1024);
    const uint64_t dB = mode == 3 ? umma_desc(smem_u32(smem) + 16384, 8192, 1024) : umma_desc(smem_u32(smem) + 16384, 16, 1024);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < REPS; ++r) {
          const uint32_t d = tmem + ((mode == 1 && (r & 1)) ? 256u : 0u);
          const uint32_t k = (uint32_t)(r & 3);
          if (mode == 2) umma_bf16_ts(d, tmem + 384u + k * 8u, dB + (mode == 3 ? k * 128u : k * 2u), idesc, r > 1);
          else umma_bf16(d, dA + k * 2u, dB + (mode == 3 ? k * 128u : k * 2u), idesc, r > 1);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, (uint32_t)pass);
      t1 = clock64();
    }
    if (threadIdx.x == 32) cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[4] = {"ss", "ss2 (two accumulators)", "ts (A in TMEM)", "ss-mn (B MN-major)"};
  const int ns[] = {16, 32, 48, 64, 96, 128, 160, 192, 224, 256};
  for (int grid = 148; grid >= 1; grid = grid == 148 ? 1 : 0) {
    for (int mode = 0; mode < 4; ++mode) {
      printf("%-26s grid=%3d :", names[mode], grid);
      for (int n : ns) {
        if (mode == 2 && n > 256) continue;
        mma_rate<<<grid, 128, 52 * 1024>>>(n, mode, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" N=%d ERR %s", n, cudaGetErrorString(e)); break; }
        long long h[148];
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < grid; ++i) c += (double)h[i];
        c /= grid;
        printf("  N=%-3d %6.1f", n, c / REPS);
      }
      printf("   clk/MMA\n");
      fflush(stdout);
    }
    if (grid == 1) break;
  }
  cudaFree(cyc);
  return 0;
}
