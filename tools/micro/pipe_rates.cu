// Pipe-rate microbenchmark for the attention softmax design (B200): warp-instructions per clock per SM sub-partition
// of the ops the exp2 split uses.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long u64;
#define ITERS 2048
#define UNROLL 16

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, float seed) {
  float f[UNROLL];
  u64 d[UNROLL];
  uint32_t h[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) {
    f[i] = seed * (threadIdx.x + i) * 1e-3f - 1.f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d[i]) : "f"(f[i]), "f"(f[i] * 0.5f));
    h[i] = 0x38003800u + i;
  }
  const u64 c1 = d[3], c2 = d[5];
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(c1), "l"(c2));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(seed), "f"(0.25f));
      if (OP == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(c1));
      if (OP == 6) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(seed), "f"(f[(i + 1) % UNROLL]));
      if (OP == 7) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(f[i]), "f"(f[(i + 1) % UNROLL]));
      if (OP == 8) { asm volatile("shl.b32 %0, %0, 23;" : "+r"(h[i])); asm volatile("add.s32 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i + 1) % UNROLL])); }
      if (OP == 9) {   // mix: 2 MUFU + 3 FFMA2 per slot pair (candidate softmax mix)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(c1), "l"(c2));
      }
      if (OP == 10) {  // mix: 1 MUFU + 1 FFMA (3-reg)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i + 7) % UNROLL]) : "f"(seed), "f"(0.25f));
      }
      if (OP == 11) {  // mix: FFMA2 + FMNMX (fma pipe + alu pipe)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(c1), "l"(c2));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed));
      }
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d[i]));
    acc += f[i] + lo + hi + __uint_as_float(h[i]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_slot) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  for (int warps = 4; warps <= 16; warps *= 2) {
    k<OP><<<148, warps * 32, 0>>>(out, cyc, 1.0001f);
    cudaDeviceSynchronize();
    k<OP><<<148, warps * 32, 0>>>(out, cyc, 1.0001f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    double winst = (double)ITERS * UNROLL * ops_per_slot * warps;    // warp-instructions per SM
    printf("%-34s warps/SM=%2d  %.3f warp-inst/clk/SM  (%.3f per sub-partition)\n", name, warps, winst / c, winst / c / 4);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.f16x2", 1);
  run<2>("ex2.approx.ftz.bf16x2", 1);
  run<3>("fma.rn.f32x2", 1);
  run<4>("fma.rn.f32", 1);
  run<5>("add.rn.f32x2", 1);
  run<6>("max.f32 (3-input)", 1);
  run<7>("cvt.rn.bf16x2.f32", 1);
  run<8>("shl+add (exponent splice)", 2);
  run<9>("mix ex2.f32 + fma.f32x2", 2);
  run<10>("mix ex2.f32 + fma.f32", 2);
  run<11>("mix fma.f32x2 + max.f32", 2);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
