// e4t_b200 — HBM-bound normalisation kernels on NHWC / token-major bf16 activations (fp32 statistics).
//   GroupNorm(32)+SiLU : diffusers ResnetBlock2D.norm1/norm2, Transformer2DModel.norm (transformer_2d.py:149,253),
//                        UNet conv_norm_out (unet_2d_condition.py:554-556)
//   LayerNorm          : BasicTransformerBlock.norm1/2/3 (attention.py:258-273)
// Coalesced 16-byte / 4-byte vector loads, warp-shuffle + shared-memory reductions, grids sized well past 148 SMs.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: sums[b][g] = (sum x, sum x^2) over HW x (C/G) elements.   (sums pre-zeroed)
// mode 0: plain stats of x.
// mode 1: backward stats: (sum dxhat, sum dxhat*xhat) with dxhat = dy * act'(y0) * gamma.
// ---------------------------------------------------------------------------------------------
static constexpr int kGNThreads = 256;
static constexpr int kGNMaxPairs = 6;  // C <= 3072

struct GNChan {
  float mean, rstd, gamma, beta;
};

__device__ __forceinline__ void gn_load_chan(GNChan* sc, const float* __restrict__ stats, const float* __restrict__ gamma,
                                              const float* __restrict__ beta, int b, int C, int G, float inv_n,
                                              float eps) {
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float s = stats[((long)b * G + g) * 2], ss = stats[((long)b * G + g) * 2 + 1];
    const float mean = s * inv_n;
    const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
    GNChan ch;
    ch.mean = mean;
    ch.rstd = rsqrtf(var + eps);
    ch.gamma = gamma[c];
    ch.beta = beta[c];
    sc[c] = ch;
  }
}

__device__ __forceinline__ float silu_grad(float y) {
  const float sg = 1.f / (1.f + __expf(-y));
  return sg * (1.f + y * (1.f - sg));
}

template <int MODE>
__global__ void __launch_bounds__(kGNThreads)
gn_stats_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ fstats,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ sums, int HW,
                int C, int G, int rows_per_cta, float eps, int act) {
  extern __shared__ __align__(16) uint8_t gsm[];
  float2* spair = reinterpret_cast<float2*>(gsm);                               // [C/2]
  GNChan* sc = reinterpret_cast<GNChan*>(gsm + (size_t)(C / 2) * sizeof(float2));  // [C] (MODE 1 only)
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(HW, r0 + rows_per_cta);
  const int npairs = C / 2;
  if (MODE == 1) {
    gn_load_chan(sc, fstats, gamma, beta, b, C, G, 1.f / ((float)HW * (float)(C / G)), eps);
    __syncthreads();
  }
  float a0[kGNMaxPairs], a1[kGNMaxPairs];
#pragma unroll
  for (int k = 0; k < kGNMaxPairs; ++k) a0[k] = a1[k] = 0.f;
  const uint32_t* xp = reinterpret_cast<const uint32_t*>(x) + ((long)b * HW) * npairs;
  const uint32_t* dp = MODE == 1 ? reinterpret_cast<const uint32_t*>(dy) + ((long)b * HW) * npairs : nullptr;
  for (int r = r0; r < r1; ++r) {
#pragma unroll
    for (int k = 0; k < kGNMaxPairs; ++k) {
      const int cp = threadIdx.x + k * kGNThreads;
      if (cp < npairs) {
        const float2 v = unpack_bf16(xp[(long)r * npairs + cp]);
        if (MODE == 0) {
          a0[k] += v.x + v.y;
          a1[k] += v.x * v.x + v.y * v.y;
        } else {
          const float2 d = unpack_bf16(dp[(long)r * npairs + cp]);
          const GNChan c0 = sc[2 * cp], c1 = sc[2 * cp + 1];
          const float xh0 = (v.x - c0.mean) * c0.rstd, xh1 = (v.y - c1.mean) * c1.rstd;
          float g0 = d.x * c0.gamma, g1 = d.y * c1.gamma;
          if (act) {
            g0 *= silu_grad(xh0 * c0.gamma + c0.beta);
            g1 *= silu_grad(xh1 * c1.gamma + c1.beta);
          }
          a0[k] += g0 + g1;
          a1[k] += g0 * xh0 + g1 * xh1;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kGNMaxPairs; ++k) {
    const int cp = threadIdx.x + k * kGNThreads;
    if (cp < npairs) spair[cp] = make_float2(a0[k], a1[k]);
  }
  __syncthreads();
  const int ppg = (C / G) / 2;  // pairs per group
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f, ss = 0.f;
    for (int i = 0; i < ppg; ++i) {
      const float2 p = spair[g * ppg + i];
      s += p.x;
      ss += p.y;
    }
    atomicAdd(&sums[((long)b * G + g) * 2], s);
    atomicAdd(&sums[((long)b * G + g) * 2 + 1], ss);
  }
}

// MODE 0: y = act(xhat*gamma+beta).   MODE 1: dx = rstd*(dxhat - mean(dxhat) - xhat*mean(dxhat*xhat)).
template <int MODE>
__global__ void __launch_bounds__(kGNThreads)
gn_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ fstats,
                const float* __restrict__ bstats, const float* __restrict__ gamma, const float* __restrict__ beta,
                bf16* __restrict__ out, int HW, int C, int G, int rows_per_cta, float eps, int act) {
  extern __shared__ __align__(16) uint8_t gsm[];
  GNChan* sc = reinterpret_cast<GNChan*>(gsm);               // [C]
  float2* sb = reinterpret_cast<float2*>(sc + C);            // [G] backward means (MODE 1)
  const int b = blockIdx.y;
  const float inv_n = 1.f / ((float)HW * (float)(C / G));
  gn_load_chan(sc, fstats, gamma, beta, b, C, G, inv_n, eps);
  if (MODE == 1) {
    for (int g = threadIdx.x; g < G; g += blockDim.x)
      sb[g] = make_float2(bstats[((long)b * G + g) * 2] * inv_n, bstats[((long)b * G + g) * 2 + 1] * inv_n);
  }
  __syncthreads();
  const int cpg = C / G;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(HW, r0 + rows_per_cta);
  const int vec_per_row = C / 8;
  const long base = ((long)b * HW + r0) * C;
  const int nvec = (r1 - r0) * vec_per_row;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int c0 = (i % vec_per_row) * 8;
    const uint4 xv = *reinterpret_cast<const uint4*>(x + base + (long)i * 8);
    uint4 dv = make_uint4(0, 0, 0, 0);
    if (MODE == 1) dv = *reinterpret_cast<const uint4*>(dy + base + (long)i * 8);
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const uint32_t ds[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t os[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = unpack_bf16(xs[j]);
      const GNChan ca = sc[c0 + 2 * j], cb = sc[c0 + 2 * j + 1];
      const float xh0 = (v.x - ca.mean) * ca.rstd, xh1 = (v.y - cb.mean) * cb.rstd;
      float o0, o1;
      if (MODE == 0) {
        o0 = xh0 * ca.gamma + ca.beta;
        o1 = xh1 * cb.gamma + cb.beta;
        if (act) {
          o0 = silu_f(o0);
          o1 = silu_f(o1);
        }
      } else {
        const float2 d = unpack_bf16(ds[j]);
        float g0 = d.x * ca.gamma, g1 = d.y * cb.gamma;
        if (act) {
          g0 *= silu_grad(xh0 * ca.gamma + ca.beta);
          g1 *= silu_grad(xh1 * cb.gamma + cb.beta);
        }
        const float2 ma = sb[(c0 + 2 * j) / cpg], mb = sb[(c0 + 2 * j + 1) / cpg];
        o0 = ca.rstd * (g0 - ma.x - xh0 * ma.y);
        o1 = cb.rstd * (g1 - mb.x - xh1 * mb.y);
      }
      os[j] = pack_bf16(o0, o1);
    }
    *reinterpret_cast<uint4*>(out + base + (long)i * 8) = make_uint4(os[0], os[1], os[2], os[3]);
  }
}

static int gn_rows_per_cta(int B, int HW) {
  // aim for >= ~8 CTAs per SM overall
  int rows = 32;
  while (rows > 1 && (long)B * cdiv(HW, rows) < 148 * 8) rows >>= 1;
  return rows;
}

// stats: fp32 [B][G][2] = (sum, sumsq); written by this call.
extern "C" int e4t_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, int B,
                                 int HW, int C, int G, float eps, int act_silu, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  E4T_CHECK(C % G == 0 && (C / G) % 2 == 0 && C % 8 == 0 && C <= 2 * kGNThreads * kGNMaxPairs,
            "e4t_groupnorm_fwd: unsupported C=%d G=%d", C, G);
  E4T_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * G * 2 * sizeof(float), st));
  const int rows = gn_rows_per_cta(B, HW);
  dim3 grid(cdiv(HW, rows), B);
  gn_stats_kernel<0><<<grid, kGNThreads, (size_t)(C / 2) * sizeof(float2), st>>>(
      (const bf16*)x, nullptr, nullptr, nullptr, nullptr, stats, HW, C, G, rows, eps, 0);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  gn_apply_kernel<0><<<grid, kGNThreads, (size_t)C * sizeof(GNChan), st>>>(
      (const bf16*)x, nullptr, stats, nullptr, gamma, beta, (bf16*)y, HW, C, G, rows, eps, act_silu);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  return 0;
}

// dx only (gamma/beta are frozen on the pre-training path).  scratch: fp32 [B][G][2].
extern "C" int e4t_groupnorm_bwd(const void* x, const void* dy, const float* gamma, const float* beta,
                                 const float* stats, void* dx, float* scratch, int B, int HW, int C, int G, float eps,
                                 int act_silu, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  E4T_CHECK(C % G == 0 && (C / G) % 2 == 0 && C % 8 == 0 && C <= 2 * kGNThreads * kGNMaxPairs,
            "e4t_groupnorm_bwd: unsupported C=%d G=%d", C, G);
  E4T_CUDA(cudaMemsetAsync(scratch, 0, (size_t)B * G * 2 * sizeof(float), st));
  static bool attr_set = false;
  if (!attr_set) {
    E4T_CUDA(cudaFuncSetAttribute(gn_stats_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(gn_apply_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  const int rows = gn_rows_per_cta(B, HW);
  dim3 grid(cdiv(HW, rows), B);
  gn_stats_kernel<1><<<grid, kGNThreads, (size_t)(C / 2) * sizeof(float2) + (size_t)C * sizeof(GNChan), st>>>(
      (const bf16*)x, (const bf16*)dy, stats, gamma, beta, scratch, HW, C, G, rows, eps, act_silu);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  gn_apply_kernel<1><<<grid, kGNThreads, (size_t)C * sizeof(GNChan) + (size_t)G * sizeof(float2), st>>>(
      (const bf16*)x, (const bf16*)dy, stats, scratch, gamma, beta, (bf16*)dx, HW, C, G, rows, eps, act_silu);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim (C <= 2048, C % 8 == 0): one warp per row, values held in registers.
// ---------------------------------------------------------------------------------------------
static constexpr int kLNMaxIter = 8;

template <int MODE>  // 0 fwd, 1 bwd(dx)
__global__ void __launch_bounds__(256)
ln_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ gamma,
          const float* __restrict__ beta, bf16* __restrict__ out, float* __restrict__ stats, long rows, int C,
          float eps) {
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = C / 8;
  float v[kLNMaxIter][8];
  const bf16* xr = x + row * C;
#pragma unroll
  for (int it = 0; it < kLNMaxIter; ++it) {
    const int vi = lane + it * 32;
    if (vi < nv) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + vi * 8);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      v[it][0] = a.x; v[it][1] = a.y; v[it][2] = b.x; v[it][3] = b.y;
      v[it][4] = c.x; v[it][5] = c.y; v[it][6] = d.x; v[it][7] = d.y;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[it][j] = 0.f;
    }
  }
  float mean, rstd;
  if (MODE == 0) {
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < kLNMaxIter; ++it)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[it][j];
    mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < kLNMaxIter; ++it) {
      const int vi = lane + it * 32;
      if (vi < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[it][j] - mean;
          q += d * d;
        }
      }
    }
    rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0) {
      stats[row * 2] = mean;
      stats[row * 2 + 1] = rstd;
    }
  } else {
    mean = stats[row * 2];
    rstd = stats[row * 2 + 1];
  }
  if (MODE == 0) {
#pragma unroll
    for (int it = 0; it < kLNMaxIter; ++it) {
      const int vi = lane + it * 32;
      if (vi < nv) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + vi * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(beta + vi * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[it][j] - mean) * rstd * gg[j] + bb[j];
        *reinterpret_cast<uint4*>(out + row * C + vi * 8) =
            make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
      }
    }
  } else {
    float dg[kLNMaxIter][8];
    float m1 = 0.f, m2 = 0.f;
    const bf16* dr = dy + row * C;
#pragma unroll
    for (int it = 0; it < kLNMaxIter; ++it) {
      const int vi = lane + it * 32;
      if (vi < nv) {
        const uint4 u = *reinterpret_cast<const uint4*>(dr + vi * 8);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        const float dd[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (v[it][j] - mean) * rstd;
          v[it][j] = xh;
          dg[it][j] = dd[j] * gg[j];
          m1 += dg[it][j];
          m2 += dg[it][j] * xh;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dg[it][j] = 0.f;
      }
    }
    m1 = warp_sum(m1) / (float)C;
    m2 = warp_sum(m2) / (float)C;
#pragma unroll
    for (int it = 0; it < kLNMaxIter; ++it) {
      const int vi = lane + it * 32;
      if (vi < nv) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dg[it][j] - m1 - v[it][j] * m2);
        *reinterpret_cast<uint4*>(out + row * C + vi * 8) =
            make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
      }
    }
  }
}

extern "C" int e4t_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats,
                                 long long rows, int C, float eps, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  E4T_CHECK(C % 8 == 0 && C <= 8 * 32 * kLNMaxIter, "e4t_layernorm_fwd: unsupported C=%d", C);
  ln_kernel<0><<<cdiv(rows, 8), 256, 0, st>>>((const bf16*)x, nullptr, gamma, beta, (bf16*)y, stats, rows, C, eps);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  return 0;
}

extern "C" int e4t_layernorm_bwd(const void* x, const void* dy, const float* gamma, const float* stats, void* dx,
                                 long long rows, int C, float eps, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  E4T_CHECK(C % 8 == 0 && C <= 8 * 32 * kLNMaxIter, "e4t_layernorm_bwd: unsupported C=%d", C);
  ln_kernel<1><<<cdiv(rows, 8), 256, 0, st>>>((const bf16*)x, (const bf16*)dy, gamma, nullptr, (bf16*)dx,
                                               const_cast<float*>(stats), rows, C, eps);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  return 0;
}
