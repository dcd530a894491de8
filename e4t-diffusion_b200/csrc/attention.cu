// e4t_b200 — fused attention core softmax(Q Kᵀ / sqrt(dh)) V for the SD-v1.4 UNet, tcgen05 + TMA (sm_100a).
// Reference call site: F.scaled_dot_product_attention in AttnProcessor2_0 (e4t/models/cross_attention.py:521-531),
// equivalently get_attention_scores + bmm (cross_attention.py:222-251,313-315).  No mask, no dropout, non-causal.
//
// Layout: Q [B][N][H*dh], K/V [B][M][H*dh] (token-major, heads are dh-wide column slices; row/batch strides are
// arguments so Q/K/V may alias one fused projection output).  dh % 8 == 0, dh <= 192.
// Head slices are addressed with 4-D tensor maps (d, head, token, batch); the d-extent of the map is dh, so TMA
// zero-fills the 64-wide smem chunk beyond dh (no padding in HBM).
//
// Forward, per CTA = one (128-query tile, head, batch):
//   warp0  TMA producer: Q once, K_j / V_j ring
//   warp1  MMA issuer  : S_j = Q·K_jᵀ into TMEM (double buffered), O += P_j·V_j (V_j as MN-major B operand)
//   warps4-7 softmax   : one query row per thread (TMEM lane == row, no shuffles): online softmax, P_j -> smem
//                        (SWIZZLE_128B, K-major A operand), lazy rescale of O in TMEM.
// Backward is two kernels that recompute P from the saved log-sum-exp:
//   dQ  kernel (CTA = query tile): S, dP = dO·Vᵀ, dS = P∘(dP − D)·scale, dQ += dS·K
//   dKV kernel (CTA = key tile)  : Sᵀ = K·Qᵀ, dPᵀ = V·dOᵀ, dV += Pᵀ·dO, dK += dSᵀ·Q
#include "common.cuh"
#include <stdlib.h>

#include "attn_common.cuh"

// =============================================================================================
// Forward
//   softmax warps are organised as CG column groups x 4 lane quadrants: warp (4 + 4*g + e) owns TMEM lanes
//   [32e, 32e+32) (the query rows) and the 16-column chunks c with (c/16) % CG == g of each S block, so that every
//   SM sub-partition has CG resident softmax warps to overlap MUFU / FMA / TMEM latencies.  Row maxima are exchanged
//   through shared memory (double buffered), row sums are combined once at the end.
// =============================================================================================
// PT (opt-in, E4T_ATTN_FWD_PT=1): P is written to TMEM (tcgen05.st, two bf16 per column) and O += P·V reads its A operand
// from there (tcgen05.mma [d], [a], bdesc): no P round trip through shared memory, and the N = dpad MMAs stop being bound
// by the 4 KiB A-tile read per instruction.
template <int CG, int OCC, bool PT>
__global__ void __launch_bounds__(128 + 128 * CG, OCC)
attn_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                const __grid_constant__ CUtensorMap mapV, const AttnArgs a) {
  constexpr int MAXC = 8 / CG;  // 16-column chunks per thread and block
  constexpr int NSOFT = 128 * CG;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps shared-space provenance
  const int q_bytes = a.DC * 16384;
  const int kv_tile = a.DC * a.BKV * 128;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + q_bytes;
  uint8_t* sV = sK + a.kst * kv_tile;
  const int p_bytes = ((a.BKV + 63) >> 6) * 16384;
  uint8_t* sP0 = sV + a.kst * kv_tile;  // pbuf x [128][BKV] bf16 (64-col chunks of 16 KiB)
  float* sMax = reinterpret_cast<float*>(sP0 + a.pbuf * p_bytes);  // [2][CG][128]
  float* sSum = sMax + 2 * CG * 128;                   // [CG][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + CG * 128);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + 2;
  uint64_t* v_full = k_empty + 2;
  uint64_t* v_empty = v_full + 2;
  uint64_t* s_full = v_empty + 2;
  uint64_t* p_ready = s_full + 2;
  uint64_t* o_done = p_ready + 1;
  uint64_t* p_free = o_done + 1;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_ready, NSOFT);
    mbar_init(o_done, 1);
    mbar_init(&p_free[0], 1);
    mbar_init(&p_free[1], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS0 = tmem, tO = tmem + (uint32_t)a.sbuf * 128u;
  const uint32_t tP0 = tO + (uint32_t)a.dpad;   // PT: pbuf x 64 columns

  // Producer and MMA warps run their control flow with all 32 lanes and issue under elect_one(): ptxas then knows a
  // single thread is active and emits the TMA / tcgen05 instructions without a per-lane retry loop.
  if (warp == 0) {
    {
      if (elect_one()) {
        mbar_expect_tx(q_full, (uint32_t)q_bytes);
        for (int c = 0; c < a.DC; ++c) tma_load_4d(sQ + c * 16384, &mapQ, q_full, c * 64, h, q0, b);
      }
      for (int j = 0; j < a.nblk; ++j) {
        const int st = j % a.kst;
        const uint32_t ph = (uint32_t)((j / a.kst) & 1);
        mbar_wait(&k_empty[st], ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&k_full[st], (uint32_t)kv_tile);
          for (int c = 0; c < a.DC; ++c)
            tma_load_4d(sK + st * kv_tile + c * a.BKV * 128, &mapK, &k_full[st], c * 64, h, j * a.BKV, b);
        }
        mbar_wait(&v_empty[st], ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&v_full[st], (uint32_t)kv_tile);
          for (int c = 0; c < a.DC; ++c)
            tma_load_4d(sV + st * kv_tile + c * a.BKV * 128, &mapV, &v_full[st], c * 64, h, j * a.BKV, b);
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc_s = umma_idesc_bf16((uint32_t)a.BKV, false, false);
      const uint32_t idesc_o = umma_idesc_bf16((uint32_t)a.dpad, false, true);
      mbar_wait(q_full, 0);
      tc_fence_after();
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
        mma_kmajor(tS0, smem_u32(sQ), 16384, smem_u32(sK), a.BKV * 128, a.dh, a.DC, idesc_s);
        umma_commit(&k_empty[0]);
        umma_commit(&s_full[0]);
      }
      for (int j = 0; j < a.nblk; ++j) {
        // sbuf == 2: S_{j+1} is issued BEFORE waiting for P_j (its TMEM buffer is the other one);
        // sbuf == 1: S_{j+1} may only overwrite the single buffer after softmax_j has read it (p_ready(j)).
        for (int pass = 0; pass < 2; ++pass) {
          const bool issue_s = (a.sbuf == 2) ? (pass == 0) : (pass == 1);
          if (issue_s && j + 1 < a.nblk) {
            const int jn = j + 1, st = jn % a.kst;
            mbar_wait(&k_full[st], (uint32_t)((jn / a.kst) & 1));
            tc_fence_after();
            if (elect_one()) {
              mma_kmajor(tS0 + (uint32_t)(jn & (a.sbuf - 1)) * 128u, smem_u32(sQ), 16384, smem_u32(sK + st * kv_tile),
                         a.BKV * 128, a.dh, a.DC, idesc_s);
              umma_commit(&k_empty[st]);
              umma_commit(&s_full[jn & (a.sbuf - 1)]);
            }
          }
          if (pass == 0) {
            const int st = j % a.kst;
            mbar_wait(p_ready, (uint32_t)(j & 1));
            mbar_wait(&v_full[st], (uint32_t)((j / a.kst) & 1));
            tc_fence_after();
            if (elect_one()) {
              if constexpr (PT)
                mma_pv_ts(tO, tP0 + (uint32_t)(j & (a.pbuf - 1)) * 64u, smem_u32(sV + st * kv_tile), a.BKV * 128, a.BKV,
                          idesc_o, j > 0 ? 1u : 0u);
              else
                mma_pv(tO, smem_u32(sP0 + (j & (a.pbuf - 1)) * p_bytes), smem_u32(sV + st * kv_tile), a.BKV * 128,
                       a.BKV, idesc_o, j > 0 ? 1u : 0u);
              umma_commit(&v_empty[st]);
              umma_commit(&p_free[j & (a.pbuf - 1)]);
              umma_commit(o_done);
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    const int ew = (warp - 4) & 3;
    const int cg = (warp - 4) >> 2;
    const int row = ew * 32 + lane;
    const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
    const float sl2 = a.scale * kLog2e;
    const int nchunk = a.BKV >> 4;
    const int ochunk = a.dpad >> 4;
    float m = -INFINITY, l = 0.f;
    const uint32_t rowoff = (uint32_t)row * 128u, r7 = (uint32_t)row & 7u;
    for (int j = 0; j < a.nblk; ++j) {
      const uint32_t tS = tS0 + (uint32_t)(j & (a.sbuf - 1)) * 128u + lane_base;
      uint8_t* sP = sP0 + (j & (a.pbuf - 1)) * p_bytes;
      const uint32_t tP = tP0 + (uint32_t)(j & (a.pbuf - 1)) * 64u + lane_base;
      mbar_wait(&s_full[j & (a.sbuf - 1)], (uint32_t)((j / a.sbuf) & 1));
      tc_fence_after();
      if (a.BKV == 128 && (j + 1) * 128 <= a.M) {
        // ======== fast path: full 128-key block; each thread owns NC chunks of 32 columns ========
        constexpr int NC = 4 / CG;
        uint32_t v[32];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          __syncwarp();
          tmem_ld32(tS + (uint32_t)((cg + i * CG) * 32), v);
          tmem_ld_wait();
          mx = max32(v, mx);
        }
        float* smx = sMax + (j & 1) * CG * 128;
        smx[cg * 128 + row] = mx;
        asm volatile("bar.sync 1, %0;" ::"n"(NSOFT) : "memory");
#pragma unroll
        for (int g = 0; g < CG; ++g) mx = fmaxf(mx, smx[g * 128 + row]);
        const float m_new = fmaxf(m, mx);
        const float alpha = ex2_approx((m - m_new) * sl2);
        // the P buffer of this block was last read by P·V of block j - pbuf
        if (j >= a.pbuf) mbar_wait(&p_free[j & (a.pbuf - 1)], (uint32_t)(((j / a.pbuf) - 1) & 1));
        float rs0 = 0.f, rs1 = 0.f;
        const float mb = m_new * sl2;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const int col0 = (cg + i * CG) * 32;
          if (NC > 1) {
            __syncwarp();
            tmem_ld32(tS + (uint32_t)col0, v);
            tmem_ld_wait();
          }
          uint32_t w[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * e]), sl2, -mb));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), sl2, -mb));
            rs0 += p0;
            rs1 += p1;
            w[e] = pack_bf16(p0, p1);
          }
          if constexpr (PT) {
            __syncwarp();
            tmem_st16(tP + (uint32_t)(col0 >> 1), w);
          } else {
            sts_row32(sP, rowoff, r7, col0, w);
          }
        }
        l = l * alpha + (rs0 + rs1);
        m = m_new;
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
          mbar_wait(o_done, (uint32_t)((j - 1) & 1));  // O is written by P·V of block j-1
          tc_fence_after();
          for (int oc = cg; oc < ochunk; oc += CG) {
            uint32_t ov[16];
            tmem_ld16(tO + lane_base + (uint32_t)(oc * 16), ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
            tmem_st16(tO + lane_base + (uint32_t)(oc * 16), ov);
          }
          tmem_st_wait();
        }
        if constexpr (PT) tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_ready);
        continue;
      }
      // ======== generic path (ragged / short key blocks) ========
      const int kv0 = j * a.BKV;
      // ---- row max over my column chunks (OCC==1: S kept in registers; OCC==2: low-register two-pass) ----
      uint32_t sv[OCC == 1 ? MAXC : 1][16];
      const bool partial = kv0 + a.BKV > a.M;  // only the last block can be ragged
      float mx = -INFINITY;
      if (OCC == 1) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
          const int ci = cg + i * CG;
          if (ci < nchunk) tmem_ld16(tS + (uint32_t)(ci * 16), sv[OCC == 1 ? i : 0]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
          const int ci = cg + i * CG;
          if (ci < nchunk) {
            uint32_t* svi = sv[OCC == 1 ? i : 0];
            if (!partial) {
#pragma unroll
              for (int e = 0; e < 16; ++e) mx = fmaxf(mx, __uint_as_float(svi[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                if (kv0 + ci * 16 + e >= a.M) svi[e] = 0xff800000u;  // -inf
                mx = fmaxf(mx, __uint_as_float(svi[e]));
              }
            }
          }
        }
      } else {
        for (int ci = cg; ci < nchunk; ci += CG) {
          __syncwarp();
          tmem_ld16(tS + (uint32_t)(ci * 16), sv[0]);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (!partial || kv0 + ci * 16 + e < a.M) mx = fmaxf(mx, __uint_as_float(sv[0][e]));
        }
      }
      // ---- exchange row maxima between the column groups ----
      float* smx = sMax + (j & 1) * CG * 128;
      smx[cg * 128 + row] = mx;
      asm volatile("bar.sync 1, %0;" ::"n"(NSOFT) : "memory");
#pragma unroll
      for (int g = 0; g < CG; ++g) mx = fmaxf(mx, smx[g * 128 + row]);
      const float m_new = fmaxf(m, mx);
      const float alpha = ex2_approx((m - m_new) * sl2);  // m = -inf on the first block -> 0
      // P buffer and O are free once PV_{j-1} retired
      if (j >= a.pbuf) mbar_wait(&p_free[j & (a.pbuf - 1)], (uint32_t)(((j / a.pbuf) - 1) & 1));
      // ---- P = exp2((s - m_new) * sl2) -> smem (bf16, swizzled K-major), partial row sum ----
      float rs = 0.f;
      const float mb = m_new * sl2;
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int ci = cg + i * CG;
        if (ci < nchunk) {
          uint32_t* svi = sv[OCC == 1 ? i : 0];
          if (OCC != 1) {
            __syncwarp();
            tmem_ld16(tS + (uint32_t)(ci * 16), svi);
            tmem_ld_wait();
            if (partial) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (kv0 + ci * 16 + e >= a.M) svi[e] = 0xff800000u;
            }
          }
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(svi[2 * e]), sl2, -mb));
            const float p1 = ex2_approx(fmaf(__uint_as_float(svi[2 * e + 1]), sl2, -mb));
            rs += p0 + p1;
            w[e] = pack_bf16(p0, p1);
          }
          if constexpr (PT) {
            __syncwarp();
            tmem_st8(tP + (uint32_t)(ci * 8), w);
          } else {
            uint8_t* pc = sP + (ci >> 2) * 16384;
            const uint32_t c16 = (uint32_t)((ci & 3) * 2);
            *reinterpret_cast<uint4*>(pc + sw128_off((uint32_t)row, c16)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(pc + sw128_off((uint32_t)row, c16 + 1)) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      }
      l = l * alpha + rs;
      m = m_new;
      // ---- lazy rescale of my chunks of the O accumulator (warp-uniform branch) ----
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        mbar_wait(o_done, (uint32_t)((j - 1) & 1));  // O is written by P·V of block j-1
        tc_fence_after();
        for (int oc = cg; oc < ochunk; oc += CG) {
          uint32_t v[16];
          tmem_ld16(tO + lane_base + (uint32_t)(oc * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
          tmem_st16(tO + lane_base + (uint32_t)(oc * 16), v);
        }
        tmem_st_wait();
      }
      if constexpr (PT) tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // ---- epilogue: combine row sums, normalise, store ----
    sSum[cg * 128 + row] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(NSOFT) : "memory");
    float lt = 0.f;
#pragma unroll
    for (int g = 0; g < CG; ++g) lt += sSum[g * 128 + row];
    mbar_wait(o_done, (uint32_t)((a.nblk - 1) & 1));
    tc_fence_after();
    const int n = q0 + row;
    const float inv_l = 1.f / lt;
    for (int oc = cg; oc < ochunk; oc += CG) {
      const int c = oc * 16;
      uint32_t v[16];
      __syncwarp();
      tmem_ld16(tO + lane_base + (uint32_t)c, v);
      tmem_ld_wait();
      if (n < a.N) {
        bf16* o = a.O + (long long)b * a.o_bs + (long long)n * a.ldo + h * a.dh + c;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          if (c + i < a.dh) {
            *reinterpret_cast<uint4*>(o + i) =
                make_uint4(pack_bf16(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l),
                           pack_bf16(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l),
                           pack_bf16(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l),
                           pack_bf16(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l));
          }
        }
      }
    }
    if (cg == 0 && n < a.N) a.LSE[((long long)b * a.H + h) * a.N + n] = m * a.scale + logf(lt);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
  }
}

// =============================================================================================
// D = rowsum(dO ∘ O)  per (b, h, n); one warp per (b, n, h)
// =============================================================================================
__global__ void attn_delta_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ Dv,
                                  int B, int H, int N, int dh, long long ldo, long long o_bs, long long lddo,
                                  long long do_bs) {
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)B * N * H) return;
  const int h = (int)(wid % H);
  const int n = (int)((wid / H) % N);
  const int b = (int)(wid / ((long long)H * N));
  const bf16* o = O + b * o_bs + (long long)n * ldo + h * dh;
  const bf16* d = dO + b * do_bs + (long long)n * lddo + h * dh;
  float acc = 0.f;
  for (int v = lane; v < dh / 8; v += 32) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + v * 8);
    const uint4 c = *reinterpret_cast<const uint4*>(d + v * 8);
    const uint32_t as[4] = {a.x, a.y, a.z, a.w}, cs[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16(as[i]), y = unpack_bf16(cs[i]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) Dv[((long long)b * H + h) * N + n] = acc;
}

// =============================================================================================
// dQ kernel: CTA = (128-query tile, head, batch); loops over key blocks.
//   TMEM: S [0,128) | dP [128,256) | dQ [256, 256+dpad)
// =============================================================================================
// Opt-in (E4T_ATTN_DELTA2=1): one THREAD per (b, n, h) instead of one warp (attn_delta_kernel keeps 27 of 32 lanes idle
// at dh = 40 and reaches 0.66 TB/s).  Consecutive threads read consecutive dh-wide slices of a token row, so the loads of a
// warp cover one contiguous span; the 4-byte results are scattered (H-strided) but are 2 % of the traffic.
__global__ void attn_delta2_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ Dv,
                                   int B, int H, int N, int dh, long long ldo, long long o_bs, long long lddo,
                                   long long do_bs) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * N * H) return;
  const int h = (int)(t % H);
  const int n = (int)((t / H) % N);
  const int b = (int)(t / ((long long)H * N));
  const bf16* o = O + b * o_bs + (long long)n * ldo + h * dh;
  const bf16* d = dO + b * do_bs + (long long)n * lddo + h * dh;
  float acc = 0.f;
  for (int v = 0; v < dh / 8; ++v) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + v * 8);
    const uint4 c = *reinterpret_cast<const uint4*>(d + v * 8);
    const uint32_t as[4] = {a.x, a.y, a.z, a.w}, cs[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16(as[i]), y = unpack_bf16(cs[i]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  Dv[((long long)b * H + h) * N + n] = acc;
}

template <int CG, int OCC>
__global__ void __launch_bounds__(128 + 128 * CG, OCC)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapdO,
                   const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps shared-space provenance
  const int q_bytes = a.DC * 16384;
  const int kv_tile = a.DC * a.BKV * 128;
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + q_bytes;
  uint8_t* sK = sdO + q_bytes;               // kst stages of {K tile, V tile}
  uint8_t* sV = sK + a.kst * kv_tile;
  uint8_t* sdS = sV + a.kst * kv_tile;  // 2 x 16 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + ((a.BKV + 63) >> 6) * 16384);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* sp_full = bars + 5;
  uint64_t* ds_ready = bars + 6;
  uint64_t* dq_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(ds_ready, 128 * CG);
    mbar_init(dq_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + (uint32_t)a.BKV, tdQ = tmem + 2u * (uint32_t)a.BKV;

  if (warp == 0) {   // whole warp runs the control flow, one elected lane issues (see attn_fwd_kernel)
    {
      if (elect_one()) {
        mbar_expect_tx(q_full, (uint32_t)(2 * q_bytes));
        for (int c = 0; c < a.DC; ++c) {
          tma_load_4d(sQ + c * 16384, &mapQ, q_full, c * 64, h, q0, b);
          tma_load_4d(sdO + c * 16384, &mapdO, q_full, c * 64, h, q0, b);
        }
      }
      for (int j = 0; j < a.nblk; ++j) {
        const int st = j % a.kst;
        mbar_wait(&kv_empty[st], (uint32_t)(((j / a.kst) & 1) ^ 1));
        if (elect_one()) {
          mbar_expect_tx(&kv_full[st], (uint32_t)(2 * kv_tile));
          for (int c = 0; c < a.DC; ++c) {
            tma_load_4d(sK + st * kv_tile + c * a.BKV * 128, &mapK, &kv_full[st], c * 64, h, j * a.BKV, b);
            tma_load_4d(sV + st * kv_tile + c * a.BKV * 128, &mapV, &kv_full[st], c * 64, h, j * a.BKV, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc_s = umma_idesc_bf16((uint32_t)a.BKV, false, false);
      const uint32_t idesc_o = umma_idesc_bf16((uint32_t)a.dpad, false, true);
      mbar_wait(q_full, 0);
      for (int j = 0; j < a.nblk; ++j) {
        const int st = j % a.kst;
        mbar_wait(&kv_full[st], (uint32_t)((j / a.kst) & 1));
        tc_fence_after();
        if (elect_one()) {
          mma_kmajor(tS, smem_u32(sQ), 16384, smem_u32(sK + st * kv_tile), a.BKV * 128, a.dh, a.DC, idesc_s);
          mma_kmajor(tdP, smem_u32(sdO), 16384, smem_u32(sV + st * kv_tile), a.BKV * 128, a.dh, a.DC, idesc_s);
          umma_commit(sp_full);
        }
        mbar_wait(ds_ready, (uint32_t)(j & 1));
        tc_fence_after();
        if (elect_one()) {
          mma_pv(tdQ, smem_u32(sdS), smem_u32(sK + st * kv_tile), a.BKV * 128, a.BKV, idesc_o, j > 0 ? 1u : 0u);
          umma_commit(&kv_empty[st]);
          umma_commit(dq_done);
        }
      }
    }
  } else if (warp >= 4) {
    constexpr int MAXC = 8 / CG;
    const int ew = (warp - 4) & 3;
    const int cg = (warp - 4) >> 2;
    const int row = ew * 32 + lane;
    const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
    const int n = q0 + row;
    const float sl2 = a.scale * kLog2e;
    const long long sidx = ((long long)b * a.H + h) * a.N + n;
    const bool n_ok = n < a.N;
    const float lse2 = n_ok ? a.LSE[sidx] * kLog2e : INFINITY;   // rows past N: p = 2^-inf = 0
    const float dlt_s = n_ok ? a.Dv[sidx] * a.scale : 0.f;
    const int nchunk = a.BKV >> 4;
    const int ochunk = a.dpad >> 4;
    const uint32_t rowoff = (uint32_t)row * 128u, r7 = (uint32_t)row & 7u;
    for (int j = 0; j < a.nblk; ++j) {
      mbar_wait(sp_full, (uint32_t)(j & 1));
      tc_fence_after();
      // sp_full(j) was committed after dQ-MMA(j-1) was issued, so the dS buffer is free here.
      if ((a.BKV & 31) == 0 && (j + 1) * a.BKV <= a.M) {
        // ======== fast path: full key block, 32-column chunks ========
        const int nc32 = a.BKV >> 5;
        for (int c = cg; c < nc32; c += CG) {
          uint32_t sreg[32], dp[32], w[16];
          __syncwarp();
          tmem_ld32(tS + lane_base + (uint32_t)(c * 32), sreg);
          tmem_ld32(tdP + lane_base + (uint32_t)(c * 32), dp);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float d0 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e]), sl2, -lse2)) *
                             fmaf(__uint_as_float(dp[2 * e]), a.scale, -dlt_s);
            const float d1 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e + 1]), sl2, -lse2)) *
                             fmaf(__uint_as_float(dp[2 * e + 1]), a.scale, -dlt_s);
            w[e] = pack_bf16(d0, d1);
          }
          sts_row32(sdS, rowoff, r7, c * 32, w);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(ds_ready);
        continue;
      }
      const int kv0 = j * a.BKV;
      const bool partial = kv0 + a.BKV > a.M;
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int ci = cg + i * CG;
        if (ci < nchunk) {
          uint32_t sreg[16], dp[16];
          __syncwarp();
          tmem_ld16(tS + lane_base + (uint32_t)(ci * 16), sreg);
          tmem_ld16(tdP + lane_base + (uint32_t)(ci * 16), dp);
          tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float d0 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e]), sl2, -lse2)) *
                       fmaf(__uint_as_float(dp[2 * e]), a.scale, -dlt_s);
            float d1 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e + 1]), sl2, -lse2)) *
                       fmaf(__uint_as_float(dp[2 * e + 1]), a.scale, -dlt_s);
            if (partial) {
              if (kv0 + ci * 16 + 2 * e >= a.M) d0 = 0.f;
              if (kv0 + ci * 16 + 2 * e + 1 >= a.M) d1 = 0.f;
            }
            w[e] = pack_bf16(d0, d1);
          }
          uint8_t* pc = sdS + (ci >> 2) * 16384;
          const uint32_t c16 = (uint32_t)((ci & 3) * 2);
          *reinterpret_cast<uint4*>(pc + sw128_off((uint32_t)row, c16)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(pc + sw128_off((uint32_t)row, c16 + 1)) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
    }
    mbar_wait(dq_done, (uint32_t)((a.nblk - 1) & 1));
    tc_fence_after();
    for (int oc = cg; oc < ochunk; oc += CG) {
      const int c = oc * 16;
      uint32_t v[16];
      __syncwarp();
      tmem_ld16(tdQ + lane_base + (uint32_t)c, v);
      tmem_ld_wait();
      if (n_ok) {
        bf16* o = a.dQ + (long long)b * a.dq_bs + (long long)n * a.lddq + h * a.dh + c;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          if (c + i < a.dh) {
            *reinterpret_cast<uint4*>(o + i) =
                make_uint4(pack_bf16(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                           pack_bf16(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                           pack_bf16(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                           pack_bf16(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
  }
}

// =============================================================================================
// dK/dV kernel: CTA = (128-key tile, head, batch); loops over query blocks of BKV (=BQ) rows.
//   TMEM: Sᵀ [0,BQ) | dPᵀ [W,W+BQ) | dV [2W,2W+dpad) | dK [2W+dpad, 2W+2*dpad)   (W = BQ_max = 64 when dpad>128 else 128)
// =============================================================================================
template <int CG, int OCC>
__global__ void __launch_bounds__(128 + 128 * CG, OCC)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                    const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapdO,
                    const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps shared-space provenance
  const int BQ = a.BKV;
  const int kv_bytes = a.DC * 16384;
  const int q_tile = a.DC * BQ * 128;
  uint8_t* sK = smem;
  uint8_t* sV = sK + kv_bytes;
  uint8_t* sQ = sV + kv_bytes;                // kst stages of {Q tile, dO tile}
  uint8_t* sdO = sQ + a.kst * q_tile;
  uint8_t* sPT = sdO + a.kst * q_tile;   // 2 x 16 KiB
  uint8_t* sdST = sPT + ((BQ + 63) >> 6) * 16384;
  float* sLSE = reinterpret_cast<float*>(sdST + ((BQ + 63) >> 6) * 16384);  // [128]
  float* sD = sLSE + 128;                                // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 128);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;   // [2]
  uint64_t* q_empty = bars + 3;  // [2]
  uint64_t* sp_full = bars + 5;
  uint64_t* ds_ready = bars + 6;
  uint64_t* acc_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(ds_ready, 128 * CG);
    mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sp_cols = (uint32_t)((BQ + 31) & ~31);
  const uint32_t tST = tmem, tdPT = tmem + sp_cols, tdV = tmem + 2 * sp_cols, tdK = tmem + 2 * sp_cols + (uint32_t)a.dpad;

  if (warp == 0) {   // whole warp runs the control flow, one elected lane issues (see attn_fwd_kernel)
    {
      if (elect_one()) {
        mbar_expect_tx(kv_full, (uint32_t)(2 * kv_bytes));
        for (int c = 0; c < a.DC; ++c) {
          tma_load_4d(sK + c * 16384, &mapK, kv_full, c * 64, h, k0, b);
          tma_load_4d(sV + c * 16384, &mapV, kv_full, c * 64, h, k0, b);
        }
      }
      for (int j = 0; j < a.nblk; ++j) {
        const int st = j % a.kst;
        mbar_wait(&q_empty[st], (uint32_t)(((j / a.kst) & 1) ^ 1));
        if (elect_one()) {
          mbar_expect_tx(&q_full[st], (uint32_t)(2 * q_tile));
          for (int c = 0; c < a.DC; ++c) {
            tma_load_4d(sQ + st * q_tile + c * BQ * 128, &mapQ, &q_full[st], c * 64, h, j * BQ, b);
            tma_load_4d(sdO + st * q_tile + c * BQ * 128, &mapdO, &q_full[st], c * 64, h, j * BQ, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc_s = umma_idesc_bf16((uint32_t)BQ, false, false);
      const uint32_t idesc_o = umma_idesc_bf16((uint32_t)a.dpad, false, true);
      mbar_wait(kv_full, 0);
      for (int j = 0; j < a.nblk; ++j) {
        const int st = j % a.kst;
        mbar_wait(&q_full[st], (uint32_t)((j / a.kst) & 1));
        tc_fence_after();
        if (elect_one()) {
          mma_kmajor(tST, smem_u32(sK), 16384, smem_u32(sQ + st * q_tile), BQ * 128, a.dh, a.DC, idesc_s);
          mma_kmajor(tdPT, smem_u32(sV), 16384, smem_u32(sdO + st * q_tile), BQ * 128, a.dh, a.DC, idesc_s);
          umma_commit(sp_full);
        }
        mbar_wait(ds_ready, (uint32_t)(j & 1));
        tc_fence_after();
        if (elect_one()) {
          mma_pv(tdV, smem_u32(sPT), smem_u32(sdO + st * q_tile), BQ * 128, BQ, idesc_o, j > 0 ? 1u : 0u);
          mma_pv(tdK, smem_u32(sdST), smem_u32(sQ + st * q_tile), BQ * 128, BQ, idesc_o, j > 0 ? 1u : 0u);
          umma_commit(&q_empty[st]);
          umma_commit(acc_done);
        }
      }
    }
  } else if (warp >= 4) {
    constexpr int MAXC = 8 / CG;
    constexpr int NSOFT = 128 * CG;
    const int ew = (warp - 4) & 3;
    const int cg = (warp - 4) >> 2;
    const int row = ew * 32 + lane;  // key index within the tile
    const int tid = threadIdx.x - 128;
    const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
    const int kv = k0 + row;
    const bool kv_ok = kv < a.M;
    const bool kv_ok_warp = (k0 + ew * 32 + 31) < a.M;   // warp-uniform
    const uint32_t rowoff = (uint32_t)row * 128u, r7 = (uint32_t)row & 7u;
    const float sl2 = a.scale * kLog2e;
    // keys past M: S^T row is 0 (zero-filled K) -> use slope 0 and rely on p = 2^(-lse) ... no: force p = 0 below
    const float sl2_row = sl2;
    const long long sbase = ((long long)b * a.H + h) * a.N;
    const int nchunk = BQ >> 4;
    const int ochunk = a.dpad >> 4;
    for (int j = 0; j < a.nblk; ++j) {
      // stage LSE / D of this query block
      const int qn = j * BQ + tid;
      if (tid < BQ) {
        sLSE[tid] = (qn < a.N) ? a.LSE[sbase + qn] * kLog2e : INFINITY;   // queries past N: p = 0
        sD[tid] = (qn < a.N) ? a.Dv[sbase + qn] * a.scale : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NSOFT) : "memory");
      mbar_wait(sp_full, (uint32_t)(j & 1));
      tc_fence_after();
      if ((BQ & 31) == 0 && kv_ok_warp) {
        // ======== fast path: 32-column chunks, whole warp inside the key range (ragged queries are handled by
        // sLSE = +inf) ========
        const int nc32 = BQ >> 5;
        for (int c32 = cg; c32 < nc32; c32 += CG) {
          const int c = c32 * 32;
          uint32_t sreg[32], dp[32], wp[16], wd[16];
          __syncwarp();
          tmem_ld32(tST + lane_base + (uint32_t)c, sreg);
          tmem_ld32(tdPT + lane_base + (uint32_t)c, dp);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float4 ls = *reinterpret_cast<const float4*>(&sLSE[c + 2 * e]);
            const float4 dd = *reinterpret_cast<const float4*>(&sD[c + 2 * e]);
            const float p0 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e]), sl2, -ls.x));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e + 1]), sl2, -ls.y));
            const float p2 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e + 2]), sl2, -ls.z));
            const float p3 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e + 3]), sl2, -ls.w));
            wp[e] = pack_bf16(p0, p1);
            wp[e + 1] = pack_bf16(p2, p3);
            wd[e] = pack_bf16(p0 * fmaf(__uint_as_float(dp[2 * e]), a.scale, -dd.x),
                              p1 * fmaf(__uint_as_float(dp[2 * e + 1]), a.scale, -dd.y));
            wd[e + 1] = pack_bf16(p2 * fmaf(__uint_as_float(dp[2 * e + 2]), a.scale, -dd.z),
                                  p3 * fmaf(__uint_as_float(dp[2 * e + 3]), a.scale, -dd.w));
          }
          sts_row32(sPT, rowoff, r7, c, wp);
          sts_row32(sdST, rowoff, r7, c, wd);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(ds_ready);
        asm volatile("bar.sync 2, %0;" ::"n"(NSOFT) : "memory");
        continue;
      }
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int ci = cg + i * CG;
        if (ci < nchunk) {
          const int c = ci * 16;
          uint32_t sreg[16], dp[16];
          __syncwarp();
          tmem_ld16(tST + lane_base + (uint32_t)c, sreg);
          tmem_ld16(tdPT + lane_base + (uint32_t)c, dp);
          tmem_ld_wait();
          uint32_t wp[8], wd[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float2 ls = *reinterpret_cast<const float2*>(&sLSE[c + 2 * e]);
            const float2 dd = *reinterpret_cast<const float2*>(&sD[c + 2 * e]);
            float p0 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e]), sl2_row, -ls.x));
            float p1 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e + 1]), sl2_row, -ls.y));
            const float d0 = p0 * fmaf(__uint_as_float(dp[2 * e]), a.scale, -dd.x);
            const float d1 = p1 * fmaf(__uint_as_float(dp[2 * e + 1]), a.scale, -dd.y);
            wp[e] = kv_ok ? pack_bf16(p0, p1) : 0u;
            wd[e] = kv_ok ? pack_bf16(d0, d1) : 0u;
          }
          const uint32_t c16 = (uint32_t)((ci & 3) * 2);
          uint8_t* pp = sPT + (ci >> 2) * 16384;
          uint8_t* pd = sdST + (ci >> 2) * 16384;
          *reinterpret_cast<uint4*>(pp + sw128_off((uint32_t)row, c16)) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
          *reinterpret_cast<uint4*>(pp + sw128_off((uint32_t)row, c16 + 1)) = make_uint4(wp[4], wp[5], wp[6], wp[7]);
          *reinterpret_cast<uint4*>(pd + sw128_off((uint32_t)row, c16)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          *reinterpret_cast<uint4*>(pd + sw128_off((uint32_t)row, c16 + 1)) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
      // everyone is done reading sLSE/sD of this block before the next block overwrites them
      asm volatile("bar.sync 2, %0;" ::"n"(NSOFT) : "memory");
    }
    mbar_wait(acc_done, (uint32_t)((a.nblk - 1) & 1));
    tc_fence_after();
    for (int which = 0; which < 2; ++which) {
      const uint32_t tacc = which == 0 ? tdV : tdK;
      bf16* base = which == 0 ? a.dV + (long long)b * a.dv_bs + (long long)kv * a.lddv
                              : a.dK + (long long)b * a.dk_bs + (long long)kv * a.lddk;
      for (int oc = cg; oc < ochunk; oc += CG) {
        const int c = oc * 16;
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tacc + lane_base + (uint32_t)c, v);
        tmem_ld_wait();
        if (kv_ok) {
          bf16* o = base + h * a.dh + c;
#pragma unroll
          for (int i = 0; i < 16; i += 8) {
            if (c + i < a.dh) {
              *reinterpret_cast<uint4*>(o + i) =
                  make_uint4(pack_bf16(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                             pack_bf16(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                             pack_bf16(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                             pack_bf16(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
  }
}

// =============================================================================================
// Fused backward (dh <= 80): CTA = (128-key tile, head, batch) looping over 128-query blocks.  S and dP are computed
// ONCE per (query block, key tile) pair (the two-kernel path recomputes them in both kernels):
//   Sᵀ = K·Qᵀ, dPᵀ = V·dOᵀ  ->  Pᵀ, dSᵀ (threads)  ->  dV += Pᵀ·dO, dK += dSᵀ·Q, dQ_blk = dS·K
// dQ_blk uses the thread-written dSᵀ tile as an MN-major A operand and is reduced into an fp32 dQ accumulator in
// HBM with red.global.add.v4.f32 (one CTA per key tile contributes to every query row).
//   TMEM: Sᵀ [0,128) | dPᵀ [128,256) | dV | dK | dQ_blk  (3 x dpad columns, dpad <= 80)
// The dQ read-out of block j-1 is deferred until after block j's dSᵀ has been handed to the tensor core.
// =============================================================================================
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mma_ds_k(uint32_t d_tmem, uint32_t sdST, uint32_t sK, uint32_t k_chunk,
                                         uint32_t idesc) {
  // D[128 q][dpad] = dS[128 q][128 kv] · K[128 kv][dpad]; A = dSᵀ tile ([kv][q], 64-q chunks of 16 KiB) read as an
  // MN-major operand, B = K tile ([kv][d], 64-wide d chunks) read as an MN-major operand.
  for (int ks = 0; ks < 8; ++ks)
    umma_bf16(d_tmem, umma_desc(sdST + ks * 2048, 16384, 1024), umma_desc(sK + ks * 2048, k_chunk, 1024), idesc,
              ks > 0 ? 1u : 0u);
}

// DQTMA (opt-in, dh <= 64): dQ_blk is staged in shared memory (fp32 [128][dh], double buffered) and reduced into the
// accumulator by warp 3 with ONE cp.reduce.async.bulk.tensor per block (full 128-byte lines through the TMA engine)
// instead of 4 red.global.add.v4.f32 per thread whose lanes each touch their own line (32 LSU wavefronts per warp
// instruction; 56 % of the kernel's LSU wavefronts in the round-1 ncu capture).
// PT (opt-in, E4T_ATTN_PT_TMEM=1): the threads write Pᵀ into 64 TMEM columns (tcgen05.st) and dV += Pᵀ·dO takes its A
// operand from TMEM: no Pᵀ stores to shared memory and no 4 KiB A-tile read per N=48 MMA (which is what makes those
// MMAs shared-memory-bandwidth bound: 5.5 KiB of operands per 24-clk instruction against 128 B/clk).
template <int CG, bool DQTMA, bool PT, bool CAUSAL = false>
__global__ void __launch_bounds__(128 + 128 * CG, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                      const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapdO,
                      const __grid_constant__ CUtensorMap mapDQ, const AttnArgs a, float* __restrict__ dQacc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int BQ = 128;
  const int kv_bytes = a.DC * 16384;
  const int q_tile = a.DC * BQ * 128;
  uint8_t* sK = smem;
  uint8_t* sV = sK + kv_bytes;
  uint8_t* sQ = sV + kv_bytes;  // kst stages of {Q tile, dO tile}
  uint8_t* sdO = sQ + a.kst * q_tile;
  uint8_t* sPT = sdO + a.kst * q_tile;  // 2 x 16 KiB
  uint8_t* sdST = sPT + 32768;          // 2 x 16 KiB
  const int dq_tile = DQTMA ? BQ * a.dh * 4 : 0;          // fp32 [128][dh] staging tile of dQ_blk (x2)
  uint8_t* sDQ = sdST + 32768;
  float* sLSE0 = reinterpret_cast<float*>(sdST + 32768 + 2 * dq_tile);  // [2 buffers][{lse [128], D [128]}], block parity
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLSE0 + 512);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;   // [2]
  uint64_t* q_empty = bars + 3;  // [2]
  uint64_t* sp_full = bars + 5;
  uint64_t* ds_ready = bars + 6;
  uint64_t* acc_done = bars + 7;
  uint64_t* dq_full = bars + 8;
  uint64_t* dq_empty = bars + 9;
  uint64_t* dq_staged = bars + 10;  // [2] DQTMA: every thread's part of the staging tile is in shared memory
  uint64_t* dq_free = bars + 12;    // [2] DQTMA: the TMA reduce has read the staging tile
  uint64_t* pt_free = bars + 14;    // reordered issue: dV(j), the reader of P^T(j) in TMEM, has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&dq_staged[i], 128 * CG);
      mbar_init(&dq_free[i], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(ds_ready, 128 * CG);
    mbar_init(acc_done, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 128 * CG);
    mbar_init(pt_free, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tST = tmem, tdPT = tmem + 128, tdV = tmem + 256, tdK = tdV + (uint32_t)a.dpad,
                 tdQ = tdK + (uint32_t)a.dpad;
  const uint32_t tPT = tdQ + (uint32_t)a.dpad;   // PT: Pᵀ as bf16 pairs, 64 columns (256 + 3*dpad + 64 <= 512)

  // Both role warps run their loops in ONE elected thread, with every shared-memory descriptor built once outside the
  // loop and the ring stage / phase tracked incrementally (round 2: the per-block `j % kst`, `j / kst` and 60 descriptor
  // constructions sat on the critical path — the thread phase and the tensor phase of a block do not overlap here, so the
  // time this thread needs to ISSUE the 30 MMAs of a block adds to every block).
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(kv_full, (uint32_t)(2 * kv_bytes));
      for (int c = 0; c < a.DC; ++c) {
        tma_load_4d(sK + c * 16384, &mapK, kv_full, c * 64, h, k0, b);
        tma_load_4d(sV + c * 16384, &mapV, kv_full, c * 64, h, k0, b);
      }
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < a.nblk; ++j) {
        mbar_wait(&q_empty[st], ph ^ 1u);
        mbar_expect_tx(&q_full[st], (uint32_t)(2 * q_tile));
        for (int c = 0; c < a.DC; ++c) {
          tma_load_4d(sQ + st * q_tile + c * BQ * 128, &mapQ, &q_full[st], c * 64, h, j * BQ, b);
          tma_load_4d(sdO + st * q_tile + c * BQ * 128, &mapdO, &q_full[st], c * 64, h, j * BQ, b);
        }
        if (++st == a.kst) {
          st = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16((uint32_t)BQ, false, false);
      const uint32_t idesc_o = umma_idesc_bf16((uint32_t)a.dpad, false, true);
      const uint32_t idesc_q = umma_idesc_bf16((uint32_t)a.dpad, true, true);
      // descriptor bases (start-address field in 16-byte units; every tile is 1024-byte aligned and all of shared memory
      // is below 256 KiB, so adding an offset never carries out of the 14-bit field)
      //   K-major tile, k-step kk of dh:            + (kk >> 2) * 1024 + (kk & 3) * 2     (64-wide chunks of 16 KiB)
      //   MN-major tile, 16-row step ks:            + ks * 128
      const uint64_t dK_k = umma_desc(smem_u32(sK), 16, 1024);          // A of S^T  = K Q^T
      const uint64_t dV_k = umma_desc(smem_u32(sV), 16, 1024);          // A of dP^T = V dO^T
      const uint64_t dK_mn = umma_desc(smem_u32(sK), 16384, 1024);      // B of dQ   = dS K
      const uint64_t dST_k = umma_desc(smem_u32(sdST), 16, 1024);       // A of dK  += dS^T Q
      const uint64_t dST_mn = umma_desc(smem_u32(sdST), 16384, 1024);   // A of dQ
      const uint64_t dPT_k = umma_desc(smem_u32(sPT), 16, 1024);        // A of dV  += P^T dO   (!PT)
      const uint64_t dQ_k0 = umma_desc(smem_u32(sQ), 16, 1024), ddO_k0 = umma_desc(smem_u32(sdO), 16, 1024);
      const uint64_t dQ_mn0 = umma_desc(smem_u32(sQ), BQ * 128, 1024), ddO_mn0 = umma_desc(smem_u32(sdO), BQ * 128, 1024);
      const uint32_t q_units = (uint32_t)q_tile >> 4;
      const int ksteps = (a.dh + 15) >> 4;
      mbar_wait(kv_full, 0);
      auto issue_s_dp = [&](uint64_t so) {      // S^T = K Q^T and dP^T = V dO^T of the block whose {Q, dO} stage starts at `so`
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t off = (uint64_t)(((kk >> 2) << 10) + ((kk & 3) << 1));
          umma_bf16(tST, dK_k + off, dQ_k0 + so + off, idesc_s, kk > 0 ? 1u : 0u);
        }
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t off = (uint64_t)(((kk >> 2) << 10) + ((kk & 3) << 1));
          umma_bf16(tdPT, dV_k + off, ddO_k0 + so + off, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(sp_full);
      };
      // REORDERED issue (PT, two {Q, dO} stages): S^T / dP^T of block j+1 are issued right after dV(j), BEFORE dK(j) and
      // dQ(j), so the threads start block j+1 after ~920 clk of tensor work instead of ~2000 and dK / dQ of block j run
      // under their arithmetic.  dS^T is double buffered for this (block parity; the second buffer is the P^T tile that
      // PT leaves unused); P^T in TMEM needs no second buffer: the threads write P^T(j+1) only after S^T(j+1) is complete,
      // which the in-order tensor pipe executes after dV(j), the reader of P^T(j).
      const bool reord = PT && a.kst == 2 && a.sbuf;
      const uint64_t dST2_k = umma_desc(smem_u32(sPT), 16, 1024), dST2_mn = umma_desc(smem_u32(sPT), 16384, 1024);
      int st = 0;
      uint32_t ph = 0;
      if (reord) {
        mbar_wait(&q_full[0], 0);
        tc_fence_after();
        issue_s_dp(0);
      }
      for (int j = 0; j < a.nblk; ++j) {
        const uint64_t so = (uint64_t)((uint32_t)st * q_units);
        const bool second = reord && (j & 1);
        const uint64_t dSk = second ? dST2_k : dST_k, dSmn = second ? dST2_mn : dST_mn;
        if (!reord) {
          mbar_wait(&q_full[st], ph);
          tc_fence_after();
          issue_s_dp(so);
        }
        mbar_wait(ds_ready, (uint32_t)(j & 1));
        tc_fence_after();
        auto next_s_dp = [&]() {
          const int stn = st ^ 1;
          mbar_wait(&q_full[stn], stn == 0 ? (ph ^ 1u) : ph);   // block j+1: next stage; the phase flips when it wraps to 0
          tc_fence_after();
          issue_s_dp((uint64_t)((uint32_t)stn * q_units));
        };
        // a.sbuf == 2: S^T / dP^T of block j+1 even before dV(j) (the threads then wait on pt_free before overwriting P^T)
        if (reord && a.sbuf == 2 && j + 1 < a.nblk) next_s_dp();
#pragma unroll
        for (int ks = 0; ks < BQ / 16; ++ks) {   // dV += P^T dO: A = P^T[:, 16 ks .. 16 ks + 16)
          const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
          if constexpr (PT) umma_bf16_ts(tdV, tPT + (uint32_t)(ks * 8), ddO_mn0 + so + (uint64_t)(ks * 128), idesc_o, acc);
          else umma_bf16(tdV, dPT_k + (uint64_t)(((ks >> 2) << 10) + ((ks & 3) << 1)), ddO_mn0 + so + (uint64_t)(ks * 128), idesc_o, acc);
        }
        if (reord) umma_commit(pt_free);         // the threads of block j+1 may now overwrite P^T
        if (reord && a.sbuf != 2 && j + 1 < a.nblk) next_s_dp();
#pragma unroll
        for (int ks = 0; ks < BQ / 16; ++ks)     // dK += dS^T Q
          umma_bf16(tdK, dSk + (uint64_t)(((ks >> 2) << 10) + ((ks & 3) << 1)), dQ_mn0 + so + (uint64_t)(ks * 128),
                    idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&q_empty[st]);               // dK was the last reader of this {Q, dO} stage (dQ reads dS^T and K only)
        if (j > 0) {
          mbar_wait(dq_empty, (uint32_t)((j - 1) & 1));  // threads have drained dQ_blk of block j-1
          tc_fence_after();
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)           // dQ_blk = dS K
          umma_bf16(tdQ, dSmn + (uint64_t)(ks * 128), dK_mn + (uint64_t)(ks * 128), idesc_q, ks > 0 ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(acc_done);
        if (++st == a.kst) {
          st = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 3) {
    if constexpr (DQTMA) {   // dQ store warp: one TMA reduce-add per query block
      for (int jb = 0; jb < a.nblk; ++jb) {
        const int buf = jb & 1;
        mbar_wait(&dq_staged[buf], (uint32_t)((jb >> 1) & 1));
        if (elect_one()) {
          tma_reduce_add_3d(&mapDQ, sDQ + buf * dq_tile, h * a.dh, jb * BQ, b);
          tma_store_commit();
          tma_store_wait_read<0>();   // the engine has read the tile: the threads may overwrite it
          mbar_arrive(&dq_free[buf]);
        }
        __syncwarp();
      }
      if (elect_one()) tma_store_wait_all();
      __syncwarp();
    }
  } else if (warp >= 4) {
    constexpr int NSOFT = 128 * CG;
    const int ew = (warp - 4) & 3;
    const int cg = (warp - 4) >> 2;
    const int row = ew * 32 + lane;  // key index within the tile == TMEM lane; also the query row of dQ_blk
    const int tid = threadIdx.x - 128;
    const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
    const int kv = k0 + row;
    const bool kv_ok = kv < a.M;
    const uint32_t rowoff = (uint32_t)row * 128u, r7 = (uint32_t)row & 7u;
    const float sl2 = a.scale * kLog2e;
    const long long sbase = ((long long)b * a.H + h) * a.N;
    const int ochunk = a.dpad >> 4;
    const int C = a.H * a.dh;

    auto drain_dq = [&](int jb) {
      // dQ_blk of block jb: TMEM lane = query row; reduce into the fp32 accumulator
      mbar_wait(dq_full, (uint32_t)(jb & 1));
      tc_fence_after();
      const int qn = jb * BQ + row;
      if constexpr (DQTMA) {
        const int buf = jb & 1;
        if (jb >= 2) mbar_wait(&dq_free[buf], (uint32_t)(((jb >> 1) - 1) & 1));  // reduce of block jb-2 has read it
        float* srow = reinterpret_cast<float*>(sDQ + buf * dq_tile) + row * a.dh;
        for (int oc = cg; oc < ochunk; oc += CG) {
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(tdQ + lane_base + (uint32_t)(oc * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            if (oc * 16 + i < a.dh)   // rows past N are clipped by the tensor map
              *reinterpret_cast<float4*>(srow + oc * 16 + i) =
                  make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                              __uint_as_float(v[i + 3]));
        }
        fence_proxy_async_smem();
        mbar_arrive(&dq_staged[buf]);
      } else {
        float* dst = dQacc + ((long long)b * a.N + qn) * C + h * a.dh;
        for (int oc = cg; oc < ochunk; oc += CG) {
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(tdQ + lane_base + (uint32_t)(oc * 16), v);
          tmem_ld_wait();
          if (qn < a.N) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              if (oc * 16 + i < a.dh)
                red_add_v4(dst + oc * 16 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                           __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(dq_empty);
    };

    // Row statistics of a query block (lse * log2e, D * scale): the first 128 threads load them from global memory ONE
    // BLOCK AHEAD into registers and publish them in a double-buffered shared-memory array.  (Round 2: loading them at
    // the top of their own block put a global-load latency in front of a 512-thread barrier in every block — 12 % of all
    // warp samples sat in that barrier, 5 % in the second barrier that protected the single buffer.)
    float lse_nx = INFINITY, d_nx = 0.f;
    if (tid < BQ && tid < a.N) {
      lse_nx = a.LSE[sbase + tid] * kLog2e;
      d_nx = a.Dv[sbase + tid] * a.scale;
    }
    for (int j = 0; j < a.nblk; ++j) {
      float* sLSE = sLSE0 + (j & 1) * 256;
      float* sD = sLSE + 128;
      if (tid < BQ) {
        sLSE[tid] = lse_nx;     // queries past N: lse = +inf -> p = 0
        sD[tid] = d_nx;
        const int qn = (j + 1) * BQ + tid;
        const bool ok = (j + 1 < a.nblk) && qn < a.N;
        lse_nx = ok ? a.LSE[sbase + qn] * kLog2e : INFINITY;
        d_nx = ok ? a.Dv[sbase + qn] * a.scale : 0.f;
      }
      // one barrier per block: a thread can only write buffer (j & 1) again in block j+2, i.e. after every thread has
      // passed the barrier of block j+1 and with it finished reading this block's values
      asm volatile("bar.sync 1, %0;" ::"n"(NSOFT) : "memory");
      mbar_wait(sp_full, (uint32_t)(j & 1));
      tc_fence_after();
      if constexpr (PT) {
        // 8-column half chunks, software pipelined: the TMEM loads of half h+1 are in flight while half h is computed
        // (one 16-column chunk at a time exposed a full tcgen05.ld latency per chunk; registers: 2 x {S, dP} x 8, as before).
        // Half hh covers columns (cg + (hh >> 1) CG) 16 + (hh & 1) 8.
        constexpr int NH = 2 * (BQ / 16) / CG;
        uint8_t* pd_tile = ((a.kst == 2 && a.sbuf && (j & 1)) ? sPT : sdST) + rowoff;
        uint32_t sA[8], dA[8], sB[8], dB[8];
        __syncwarp();
        tmem_ld8(tST + lane_base + (uint32_t)(cg * 16), sA);
        tmem_ld8(tdPT + lane_base + (uint32_t)(cg * 16), dA);
        tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          const int c = (cg + (hh >> 1) * CG) * 16 + (hh & 1) * 8;
          if (hh + 1 < NH) {
            const int cn = (cg + ((hh + 1) >> 1) * CG) * 16 + ((hh + 1) & 1) * 8;
            __syncwarp();
            tmem_ld8(tST + lane_base + (uint32_t)cn, (hh & 1) ? sA : sB);
            tmem_ld8(tdPT + lane_base + (uint32_t)cn, (hh & 1) ? dA : dB);
          }
          const uint32_t* sr = (hh & 1) ? sB : sA;
          const uint32_t* dr = (hh & 1) ? dB : dA;
          uint32_t wp[4], wd[4];
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const float4 ls = *reinterpret_cast<const float4*>(&sLSE[c + 2 * e]);
            const float4 dd = *reinterpret_cast<const float4*>(&sD[c + 2 * e]);
            float p0 = ex2_approx(fmaf(__uint_as_float(sr[2 * e]), sl2, -ls.x));
            float p1 = ex2_approx(fmaf(__uint_as_float(sr[2 * e + 1]), sl2, -ls.y));
            float p2 = ex2_approx(fmaf(__uint_as_float(sr[2 * e + 2]), sl2, -ls.z));
            float p3 = ex2_approx(fmaf(__uint_as_float(sr[2 * e + 3]), sl2, -ls.w));
            if constexpr (CAUSAL) {   // key kv attends only to queries >= kv: P and dS of earlier queries are exactly 0
              const int qb = j * BQ + c + 2 * e;
              p0 = kv > qb ? 0.f : p0;
              p1 = kv > qb + 1 ? 0.f : p1;
              p2 = kv > qb + 2 ? 0.f : p2;
              p3 = kv > qb + 3 ? 0.f : p3;
            }
            wp[e] = pack_bf16(p0, p1);
            wp[e + 1] = pack_bf16(p2, p3);
            wd[e] = pack_bf16(p0 * fmaf(__uint_as_float(dr[2 * e]), a.scale, -dd.x),
                              p1 * fmaf(__uint_as_float(dr[2 * e + 1]), a.scale, -dd.y));
            wd[e + 1] = pack_bf16(p2 * fmaf(__uint_as_float(dr[2 * e + 2]), a.scale, -dd.z),
                                  p3 * fmaf(__uint_as_float(dr[2 * e + 3]), a.scale, -dd.w));
          }
          if (!kv_ok) {   // keys past M (only in the last key tile): contribute nothing
#pragma unroll
            for (int e = 0; e < 4; ++e) wp[e] = wd[e] = 0u;
          }
          if (hh == 0 && a.kst == 2 && a.sbuf && j > 0) {   // reordered issue: dV(j-1) still reads P^T(j-1) after S/dP(j)
            mbar_wait(pt_free, (uint32_t)((j - 1) & 1));
            tc_fence_after();
          }
          __syncwarp();
          tmem_st4(tPT + lane_base + (uint32_t)(c >> 1), wp);   // 8 bf16 of this key row -> 4 columns
          const uint32_t cb = (uint32_t)((c & 63) >> 3);
          *reinterpret_cast<uint4*>(pd_tile + (c >> 6) * 16384 + ((cb ^ r7) << 4)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          if (hh + 1 < NH) tmem_ld_wait();
        }
      } else {
      // 16-column chunks (two TMEM loads in flight per chunk); chunk c16 belongs to column group (c16 % CG)
      for (int c16 = cg; c16 < BQ / 16; c16 += CG) {
        const int c = c16 * 16;
        uint32_t sreg[16], dp[16], wp[8], wd[8];
        __syncwarp();
        tmem_ld16(tST + lane_base + (uint32_t)c, sreg);
        tmem_ld16(tdPT + lane_base + (uint32_t)c, dp);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float4 ls = *reinterpret_cast<const float4*>(&sLSE[c + 2 * e]);
          const float4 dd = *reinterpret_cast<const float4*>(&sD[c + 2 * e]);
          float p0 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e]), sl2, -ls.x));
          float p1 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e + 1]), sl2, -ls.y));
          float p2 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e + 2]), sl2, -ls.z));
          float p3 = ex2_approx(fmaf(__uint_as_float(sreg[2 * e + 3]), sl2, -ls.w));
          if constexpr (CAUSAL) {
            const int qb = j * BQ + c + 2 * e;
            p0 = kv > qb ? 0.f : p0;
            p1 = kv > qb + 1 ? 0.f : p1;
            p2 = kv > qb + 2 ? 0.f : p2;
            p3 = kv > qb + 3 ? 0.f : p3;
          }
          wp[e] = pack_bf16(p0, p1);
          wp[e + 1] = pack_bf16(p2, p3);
          wd[e] = pack_bf16(p0 * fmaf(__uint_as_float(dp[2 * e]), a.scale, -dd.x),
                            p1 * fmaf(__uint_as_float(dp[2 * e + 1]), a.scale, -dd.y));
          wd[e + 1] = pack_bf16(p2 * fmaf(__uint_as_float(dp[2 * e + 2]), a.scale, -dd.z),
                                p3 * fmaf(__uint_as_float(dp[2 * e + 3]), a.scale, -dd.w));
        }
        if (!kv_ok) {   // keys past M (only in the last key tile): contribute nothing
#pragma unroll
          for (int e = 0; e < 8; ++e) wp[e] = wd[e] = 0u;
        }
        uint8_t* pd = sdST + (c >> 6) * 16384 + rowoff;
        const uint32_t cb = (uint32_t)((c & 63) >> 3);
        if constexpr (PT) {
          if (a.kst == 2 && a.sbuf && j > 0 && c16 == cg) {   // reordered issue: dV(j-1) still reads P^T(j-1) after S/dP(j)
            mbar_wait(pt_free, (uint32_t)((j - 1) & 1));
            tc_fence_after();
          }
          __syncwarp();
          tmem_st8(tPT + lane_base + (uint32_t)(c >> 1), wp);   // 16 bf16 of this key row -> 8 columns
        } else {
          uint8_t* pp = sPT + (c >> 6) * 16384 + rowoff;
          *reinterpret_cast<uint4*>(pp + ((cb ^ r7) << 4)) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
          *reinterpret_cast<uint4*>(pp + (((cb + 1) ^ r7) << 4)) = make_uint4(wp[4], wp[5], wp[6], wp[7]);
        }
        *reinterpret_cast<uint4*>(pd + ((cb ^ r7) << 4)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        *reinterpret_cast<uint4*>(pd + (((cb + 1) ^ r7) << 4)) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
      }
      }
      if constexpr (PT) tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
      if (j > 0) drain_dq(j - 1);
    }
    drain_dq(a.nblk - 1);
    mbar_wait(acc_done, (uint32_t)((a.nblk - 1) & 1));
    tc_fence_after();
    for (int which = 0; which < 2; ++which) {
      const uint32_t tacc = which == 0 ? tdV : tdK;
      bf16* base = which == 0 ? a.dV + (long long)b * a.dv_bs + (long long)kv * a.lddv
                              : a.dK + (long long)b * a.dk_bs + (long long)kv * a.lddk;
      for (int oc = cg; oc < ochunk; oc += CG) {
        const int c = oc * 16;
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tacc + lane_base + (uint32_t)c, v);
        tmem_ld_wait();
        if (kv_ok) {
          bf16* o = base + h * a.dh + c;
#pragma unroll
          for (int i = 0; i < 16; i += 8) {
            if (c + i < a.dh) {
              *reinterpret_cast<uint4*>(o + i) =
                  make_uint4(pack_bf16(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                             pack_bf16(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                             pack_bf16(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                             pack_bf16(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

__global__ void cvt_dq_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long rows_per_b, int C,
                              long long ldd, long long d_bs, long long total_vec) {
  const int vpr = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c = (int)(i % vpr) * 8;
    const float4 x = *reinterpret_cast<const float4*>(src + r * C + c);
    const float4 y = *reinterpret_cast<const float4*>(src + r * C + c + 4);
    const long long bb = r / rows_per_b, rr = r % rows_per_b;
    *reinterpret_cast<uint4*>(dst + bb * d_bs + rr * ldd + c) =
        make_uint4(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w), pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
  }
}

// =============================================================================================
// Host
// =============================================================================================
static int make_head_map(CUtensorMap* m, const void* p, int dh, int H, int rows, int B, long long ld, long long bs,
                         int box_rows) {
  uint64_t dims[4] = {(uint64_t)dh, (uint64_t)H, (uint64_t)rows, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)dh * 2, (uint64_t)ld * 2, (uint64_t)(B > 1 ? bs : (long long)rows * ld) * 2};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return e4t_tmap_encode(m, p, 4, dims, str, box, 2);
}
int e4t_attn_make_head_map(CUtensorMap* m, const void* p, int dh, int H, int rows, int B, long long ld, long long bs,
                           int box_rows) {
  return make_head_map(m, p, dh, H, rows, B, ld, bs, box_rows);
}
// two-tile ping-pong forward (attention_fwd2.cu): 1 = launched, 0 = shape outside its envelope, < 0 = error
int e4t_attn_fwd2_try(const void* Q, const void* K, const void* V, void* O, float* LSE, int B, int H, int N, int M, int dh,
                      long long ldq, long long q_bs, long long ldk, long long k_bs, long long ldv, long long v_bs,
                      long long ldo, long long o_bs, float scale, cudaStream_t st);
static int round16(int x) { return (x + 15) / 16 * 16; }
// column groups of softmax/dS warps per kernel (0 fwd, 1 dQ, 2 dKV); E4T_ATTN_CG="f,q,k" overrides for tuning
static int attn_cg(int which) {
  // defaults 4,4,4 and two-CTAs-per-SM variants allowed; the environment is parsed on every call (cheap) so a tuning
  // script can sweep the variants in one process
  int cfg[4] = {4, 4, 4, 2};   // 4th: 0 never two CTAs/SM, 1 long key axes only, 2 also single-key-block shapes (0.132 -> 0.083 ms
                               // at the level-0 cross-attention shape, r02 call 17)
  const char* e = getenv("E4T_ATTN_CG");
  if (e) sscanf(e, "%d,%d,%d,%d", &cfg[0], &cfg[1], &cfg[2], &cfg[3]);
  for (int i = 0; i < 3; ++i) if (cfg[i] != 2 && cfg[i] != 4) cfg[i] = 4;
  return cfg[which];
}

static void launch_attn_delta(const void* O, const void* dO, float* Dv, int B, int H, int N, int dh, long long ldo,
                              long long o_bs, long long lddo, long long do_bs, cudaStream_t st) {
  const char* e = getenv("E4T_ATTN_DELTA2");   // default ON since round 2 (thread-per-row; =0 selects the warp-per-row kernel)
  if (!e || atoi(e) != 0)
    attn_delta2_kernel<<<cdiv((long long)B * N * H, 256), 256, 0, st>>>((const bf16*)O, (const bf16*)dO, Dv, B, H, N, dh,
                                                                        ldo, o_bs, lddo, do_bs);
  else
    attn_delta_kernel<<<cdiv((long long)B * N * H, 8), 256, 0, st>>>((const bf16*)O, (const bf16*)dO, Dv, B, H, N, dh,
                                                                     ldo, o_bs, lddo, do_bs);
}

static int attn_common_checks(int dh, long long ldq, long long ldk, long long ldv) {
  E4T_CHECK(dh % 8 == 0 && dh >= 8 && dh <= 192, "attention: head dim %d unsupported (need dh %% 8 == 0, <= 192)", dh);
  E4T_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0, "attention: row strides must be multiples of 8 elements");
  return 0;
}

extern "C" int e4t_attn_fwd(const void* Q, const void* K, const void* V, void* O, float* LSE, int B, int H, int N,
                            int M, int dh, long long ldq, long long q_bs, long long ldk, long long k_bs,
                            long long ldv, long long v_bs, long long ldo, long long o_bs, float scale,
                            void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int e = attn_common_checks(dh, ldq, ldk, ldv)) return e;
  {
    const int r = e4t_attn_fwd2_try(Q, K, V, O, LSE, B, H, N, M, dh, ldq, q_bs, ldk, k_bs, ldv, v_bs, ldo, o_bs, scale, st);
    if (r < 0) return e4t_set_error("e4t_attn_fwd: two-tile kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (r == 1) {
      E4T_COUNT_LAUNCH();
      return 0;
    }
  }
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.N = N; a.M = M; a.dh = dh;
  a.DC = cdiv(dh, 64);
  a.dpad = round16(dh);
  a.BKV = M >= 128 ? 128 : round16(M);
  a.nblk = cdiv(M, a.BKV);
  // dh <= 64 and a long key axis: two CTAs per SM (256 TMEM columns, single S buffer, single K/V stage)
  // ... or a single key block (cross-attention, M <= 128; E4T_ATTN_CG 4th field >= 2): nothing to pipeline inside the CTA, so a
  // second resident CTA hides the fixed per-CTA latencies (TMEM alloc, barrier init, the one TMA -> MMA -> softmax -> MMA chain)
  const int occ_mode = attn_cg(3);
  const bool occ2 = a.DC == 1 && occ_mode != 0 && (a.nblk >= 4 || (occ_mode >= 2 && a.nblk == 1));
  a.kst = (a.DC >= 3 || occ2) ? 1 : 2;
  a.sbuf = occ2 ? 1 : 2;
  a.tmem_cols = occ2 ? 256 : 512;
  a.pbuf = 1;
  a.scale = scale;
  a.O = (bf16*)O; a.ldo = ldo; a.o_bs = o_bs; a.LSE = LSE;
  CUtensorMap mQ, mK, mV;
  if (int e = make_head_map(&mQ, Q, dh, H, N, B, ldq, q_bs, 128)) return e;
  if (int e = make_head_map(&mK, K, dh, H, M, B, ldk, k_bs, a.BKV)) return e;
  if (int e = make_head_map(&mV, V, dh, H, M, B, ldv, v_bs, a.BKV)) return e;
  const int cg = attn_cg(0);
  const size_t smem_base = (size_t)a.DC * 16384 + (size_t)2 * a.kst * a.DC * a.BKV * 128 + 3 * 4 * 128 * 4 + 256 + 1024;
  const size_t p_bytes = (size_t)cdiv(a.BKV, 64) * 16384;
  if (!occ2 && a.nblk > 1 && smem_base + 2 * p_bytes <= 227 * 1024) a.pbuf = 2;
  const size_t smem = smem_base + a.pbuf * p_bytes;
  static bool attr = false;
  if (!attr) {
    E4T_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<2, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<4, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<2, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<2, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<4, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    attr = true;
  }
  E4T_CHECK(smem <= (occ2 ? 113 : 227) * 1024, "e4t_attn_fwd: smem budget exceeded (%zu)", smem);
  // opt-in: P through TMEM (needs sbuf*128 + dpad + pbuf*64 columns)
  const char* pte = getenv("E4T_ATTN_FWD_PT");
  const bool pt = (!pte || atoi(pte) != 0) && a.sbuf * 128 + a.dpad + a.pbuf * 64 <= a.tmem_cols;   // default ON
  const dim3 grid(cdiv(N, 128), H, B);
  if (pt) {
    if (occ2) attn_fwd_kernel<2, 2, true><<<grid, 128 + 128 * 2, smem, st>>>(mQ, mK, mV, a);
    else if (cg == 4) attn_fwd_kernel<4, 1, true><<<grid, 128 + 128 * 4, smem, st>>>(mQ, mK, mV, a);
    else attn_fwd_kernel<2, 1, true><<<grid, 128 + 128 * 2, smem, st>>>(mQ, mK, mV, a);
  } else if (occ2) attn_fwd_kernel<2, 2, false><<<grid, 128 + 128 * 2, smem, st>>>(mQ, mK, mV, a);
  else if (cg == 4) attn_fwd_kernel<4, 1, false><<<grid, 128 + 128 * 4, smem, st>>>(mQ, mK, mV, a);
  else attn_fwd_kernel<2, 1, false><<<grid, 128 + 128 * 2, smem, st>>>(mQ, mK, mV, a);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  return 0;
}

// Dv: fp32 scratch [B][H][N].
extern "C" int e4t_attn_bwd(const void* Q, const void* K, const void* V, const void* O, const void* dO,
                            const float* LSE, float* Dv, void* dQ, void* dK, void* dV, int B, int H, int N, int M,
                            int dh, long long ldq, long long q_bs, long long ldk, long long k_bs, long long ldv,
                            long long v_bs, long long ldo, long long o_bs, long long lddo, long long do_bs,
                            long long lddq, long long dq_bs, long long lddk, long long dk_bs, long long lddv,
                            long long dv_bs, float scale, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int e = attn_common_checks(dh, ldq, ldk, ldv)) return e;
  E4T_CHECK(lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0, "e4t_attn_bwd: strides %% 8");
  launch_attn_delta(O, dO, Dv, B, H, N, dh, ldo, o_bs, lddo, do_bs, st);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  static bool attr = false;
  if (!attr) {
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    attr = true;
  }
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.N = N; a.M = M; a.dh = dh;
  a.DC = cdiv(dh, 64);
  a.dpad = round16(dh);
  a.scale = scale;
  a.LSE = const_cast<float*>(LSE); a.Dv = Dv;
  a.dQ = (bf16*)dQ; a.lddq = lddq; a.dq_bs = dq_bs;
  a.dK = (bf16*)dK; a.lddk = lddk; a.dk_bs = dk_bs;
  a.dV = (bf16*)dV; a.lddv = lddv; a.dv_bs = dv_bs;
  {  // dQ
    const bool occ2 = (a.DC == 1 && M >= 512 && attn_cg(3) != 0);   // two CTAs/SM: 64-wide key blocks, 256 TMEM cols
    a.BKV = occ2 ? 64 : (M >= 128 ? 128 : round16(M));
    a.nblk = cdiv(M, a.BKV);
    a.tmem_cols = occ2 ? 256 : 512;
    CUtensorMap mQ, mK, mV, mdO;
    if (int e = make_head_map(&mQ, Q, dh, H, N, B, ldq, q_bs, 128)) return e;
    if (int e = make_head_map(&mdO, dO, dh, H, N, B, lddo, do_bs, 128)) return e;
    if (int e = make_head_map(&mK, K, dh, H, M, B, ldk, k_bs, a.BKV)) return e;
    if (int e = make_head_map(&mV, V, dh, H, M, B, ldv, v_bs, a.BKV)) return e;
    const size_t fixed = (size_t)2 * a.DC * 16384 + (size_t)cdiv(a.BKV, 64) * 16384 + 256 + 1024;
    const size_t per_stage = (size_t)2 * a.DC * a.BKV * 128;
    const size_t cap = (size_t)(occ2 ? 113 : 227) * 1024;
    a.kst = (fixed + 2 * per_stage <= cap && a.nblk > 1) ? 2 : 1;
    const size_t smem = fixed + a.kst * per_stage;
    E4T_CHECK(smem <= cap, "e4t_attn_bwd(dQ): smem budget exceeded (%zu)", smem);
    if (occ2) attn_bwd_dq_kernel<2, 2><<<dim3(cdiv(N, 128), H, B), 128 + 128 * 2, smem, st>>>(mQ, mK, mV, mdO, a);
    else if (attn_cg(1) == 4) attn_bwd_dq_kernel<4, 1><<<dim3(cdiv(N, 128), H, B), 128 + 128 * 4, smem, st>>>(mQ, mK, mV, mdO, a);
    else attn_bwd_dq_kernel<2, 1><<<dim3(cdiv(N, 128), H, B), 128 + 128 * 2, smem, st>>>(mQ, mK, mV, mdO, a);
    E4T_COUNT_LAUNCH();
    E4T_LAUNCH_CHECK();
  }
  {  // dK, dV
    const bool occ2 = (a.DC == 1 && N >= 512 && attn_cg(3) != 0);   // two CTAs/SM: 64-wide query blocks
    const int bq_max = (a.dpad > 128 || occ2) ? 64 : 128;
    a.BKV = N >= bq_max ? bq_max : round16(N);
    a.nblk = cdiv(N, a.BKV);
    a.tmem_cols = occ2 ? 256 : 512;
    CUtensorMap mQ, mK, mV, mdO;
    if (int e = make_head_map(&mQ, Q, dh, H, N, B, ldq, q_bs, a.BKV)) return e;
    if (int e = make_head_map(&mdO, dO, dh, H, N, B, lddo, do_bs, a.BKV)) return e;
    if (int e = make_head_map(&mK, K, dh, H, M, B, ldk, k_bs, 128)) return e;
    if (int e = make_head_map(&mV, V, dh, H, M, B, ldv, v_bs, 128)) return e;
    const size_t fixed = (size_t)2 * a.DC * 16384 + (size_t)2 * cdiv(a.BKV, 64) * 16384 + 1024 + 256 + 1024;
    const size_t per_stage = (size_t)2 * a.DC * a.BKV * 128;
    const size_t cap = (size_t)(occ2 ? 113 : 227) * 1024;
    a.kst = (fixed + 2 * per_stage <= cap && a.nblk > 1) ? 2 : 1;
    const size_t smem = fixed + a.kst * per_stage;
    E4T_CHECK(smem <= cap, "e4t_attn_bwd(dKV): smem budget exceeded (%zu)", smem);
    if (occ2) attn_bwd_dkv_kernel<2, 2><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 2, smem, st>>>(mQ, mK, mV, mdO, a);
    else if (attn_cg(2) == 4) attn_bwd_dkv_kernel<4, 1><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 4, smem, st>>>(mQ, mK, mV, mdO, a);
    else attn_bwd_dkv_kernel<2, 1><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 2, smem, st>>>(mQ, mK, mV, mdO, a);
    E4T_COUNT_LAUNCH();
    E4T_LAUNCH_CHECK();
  }
  return 0;
}

static int attn_bwd_fused_impl(const void* Q, const void* K, const void* V, const void* O, const void* dO,
                               const float* LSE, float* Dv, float* dQacc, void* dQ, void* dK, void* dV, int B, int H,
                               int N, int M, int dh, long long ldq, long long q_bs, long long ldk, long long k_bs,
                               long long ldv, long long v_bs, long long ldo, long long o_bs, long long lddo,
                               long long do_bs, long long lddq, long long dq_bs, long long lddk, long long dk_bs,
                               long long lddv, long long dv_bs, float scale, int causal, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (int e = attn_common_checks(dh, ldq, ldk, ldv)) return e;
  const int dpad = round16(dh);
  E4T_CHECK(!causal || (256 + 3 * dpad <= 512 && N == M), "e4t_attn_bwd_fused_causal: needs dh <= 80 and N == M");
  if (!causal && (256 + 3 * dpad > 512 || N < 128))
    return e4t_attn_bwd(Q, K, V, O, dO, LSE, Dv, dQ, dK, dV, B, H, N, M, dh, ldq, q_bs, ldk, k_bs, ldv, v_bs, ldo, o_bs,
                        lddo, do_bs, lddq, dq_bs, lddk, dk_bs, lddv, dv_bs, scale, stream_);
  E4T_CHECK(lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0, "e4t_attn_bwd_fused: strides %% 8");
  launch_attn_delta(O, dO, Dv, B, H, N, dh, ldo, o_bs, lddo, do_bs, st);
  E4T_COUNT_LAUNCH();
  const long long nacc = (long long)B * N * H * dh;
  E4T_CUDA(cudaMemsetAsync(dQacc, 0, (size_t)nacc * sizeof(float), st));
  static bool attr = false;
  if (!attr) {
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<4, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<4, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    E4T_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<4, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.N = N; a.M = M; a.dh = dh;
  a.DC = cdiv(dh, 64);
  a.dpad = dpad;
  a.scale = scale;
  a.LSE = const_cast<float*>(LSE); a.Dv = Dv;
  a.dK = (bf16*)dK; a.lddk = lddk; a.dk_bs = dk_bs;
  a.dV = (bf16*)dV; a.lddv = lddv; a.dv_bs = dv_bs;
  a.BKV = 128;
  a.nblk = cdiv(N, 128);
  CUtensorMap mQ, mK, mV, mdO;
  if (int e = make_head_map(&mQ, Q, dh, H, N, B, ldq, q_bs, 128)) return e;
  if (int e = make_head_map(&mdO, dO, dh, H, N, B, lddo, do_bs, 128)) return e;
  if (int e = make_head_map(&mK, K, dh, H, M, B, ldk, k_bs, 128)) return e;
  if (int e = make_head_map(&mV, V, dh, H, M, B, ldv, v_bs, 128)) return e;
  const size_t fixed = (size_t)2 * a.DC * 16384 + 65536 + 2048 + 256 + 1024;   // K, V | P^T, dS^T | 2 x {lse, D} | barriers | align
  const size_t per_stage = (size_t)2 * a.DC * 128 * 128;
  a.kst = (fixed + 2 * per_stage <= 227 * 1024 && a.nblk > 1) ? 2 : 1;
  const size_t smem = fixed + a.kst * per_stage;
  E4T_CHECK(smem <= 227 * 1024, "e4t_attn_bwd_fused: smem budget exceeded (%zu)", smem);
  // opt-in: dQ through a TMA reduce-add (needs 2 x 128 x dh fp32 of extra shared memory: dh <= 64 only)
  const char* dqe = getenv("E4T_ATTN_DQ_TMA");
  const bool dq_tma = a.DC == 1 && (!dqe || atoi(dqe) != 0) && smem + (size_t)2 * 128 * dh * 4 <= 227 * 1024;   // default ON
  // opt-in: Pᵀ through TMEM (256 + 3*dpad + 64 columns must fit 512: dpad <= 64)
  const char* pte = getenv("E4T_ATTN_PT_TMEM");
  const bool pt_tmem = (!pte || atoi(pte) != 0) && 256 + 3 * dpad + 64 <= 512;   // default ON
  {   // (field reused) 1 = reordered MMA issue with double-buffered dS^T (default; E4T_ATTN_BWD_REORD=0 keeps the block order)
    const char* re = getenv("E4T_ATTN_BWD_REORD");
    a.sbuf = re ? atoi(re) : 1;       // 0 block order, 1 dV(j) | S,dP(j+1) | dK,dQ(j), 2 S,dP(j+1) | dV(j) | dK,dQ(j)
  }
  CUtensorMap mDQ;
  memset(&mDQ, 0, sizeof(mDQ));
  if (dq_tma) {
    const uint64_t dims[3] = {(uint64_t)H * dh, (uint64_t)N, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)H * dh * 4, (uint64_t)N * H * dh * 4};
    const uint32_t box[3] = {(uint32_t)dh, 128, 1};
    if (int e = e4t_tmap_encode(&mDQ, dQacc, 3, dims, str, box, 4, 0)) return e;
  }
  a.causal = causal;
  if (causal && dq_tma && pt_tmem) {
    attn_bwd_fused_kernel<4, true, true, true><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 4, smem + 2 * 128 * dh * 4, st>>>(
        mQ, mK, mV, mdO, mDQ, a, dQacc);
  } else if (causal) {
    attn_bwd_fused_kernel<4, false, false, true><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 4, smem, st>>>(mQ, mK, mV, mdO, mDQ,
                                                                                                         a, dQacc);
  } else if (dq_tma && pt_tmem) {
    attn_bwd_fused_kernel<4, true, true><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 4, smem + 2 * 128 * dh * 4, st>>>(
        mQ, mK, mV, mdO, mDQ, a, dQacc);
  } else if (dq_tma) {
    attn_bwd_fused_kernel<4, true, false><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 4, smem + 2 * 128 * dh * 4, st>>>(
        mQ, mK, mV, mdO, mDQ, a, dQacc);
  } else if (pt_tmem) {
    attn_bwd_fused_kernel<4, false, true><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 4, smem, st>>>(mQ, mK, mV, mdO, mDQ, a,
                                                                                                  dQacc);
  } else if (attn_cg(1) == 4) attn_bwd_fused_kernel<4, false, false><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 4, smem, st>>>(mQ, mK, mV, mdO, mDQ, a, dQacc);
  else attn_bwd_fused_kernel<2, false, false><<<dim3(cdiv(M, 128), H, B), 128 + 128 * 2, smem, st>>>(mQ, mK, mV, mdO, mDQ, a, dQacc);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  const long long total_vec = (long long)B * N * (H * dh / 8);
  long long blocks = (total_vec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cvt_dq_kernel<<<(int)blocks, 256, 0, st>>>(dQacc, (bf16*)dQ, N, H * dh, lddq, dq_bs, total_vec);
  E4T_COUNT_LAUNCH();
  E4T_LAUNCH_CHECK();
  return 0;
}

// Fused backward (one pass over the (query block, key tile) pairs).  dQacc: fp32 scratch [B][N][H*dh] (zeroed here).
// Falls back to the two-kernel path when the head dim does not fit the TMEM budget (dh > 80) or N is tiny.
extern "C" int e4t_attn_bwd_fused(const void* Q, const void* K, const void* V, const void* O, const void* dO,
                                  const float* LSE, float* Dv, float* dQacc, void* dQ, void* dK, void* dV, int B, int H,
                                  int N, int M, int dh, long long ldq, long long q_bs, long long ldk, long long k_bs,
                                  long long ldv, long long v_bs, long long ldo, long long o_bs, long long lddo,
                                  long long do_bs, long long lddq, long long dq_bs, long long lddk, long long dk_bs,
                                  long long lddv, long long dv_bs, float scale, void* stream_) {
  return attn_bwd_fused_impl(Q, K, V, O, dO, LSE, Dv, dQacc, dQ, dK, dV, B, H, N, M, dh, ldq, q_bs, ldk, k_bs, ldv, v_bs, ldo,
                             o_bs, lddo, do_bs, lddq, dq_bs, lddk, dk_bs, lddv, dv_bs, scale, 0, stream_);
}
// The same with a causal mask (key j contributes to query i only if j <= i; N == M, dh <= 80, any N): the backward of the
// CLIP text tower's self-attention (e4t/models/modeling_clip.py:45-51) on the tensor cores.  O and LSE come from the
// forward that applied the same mask (e4t_attn_small_fwd).
extern "C" int e4t_attn_bwd_fused_causal(const void* Q, const void* K, const void* V, const void* O, const void* dO,
                                         const float* LSE, float* Dv, float* dQacc, void* dQ, void* dK, void* dV, int B,
                                         int H, int N, int M, int dh, long long ldq, long long q_bs, long long ldk,
                                         long long k_bs, long long ldv, long long v_bs, long long ldo, long long o_bs,
                                         long long lddo, long long do_bs, long long lddq, long long dq_bs, long long lddk,
                                         long long dk_bs, long long lddv, long long dv_bs, float scale, void* stream_) {
  return attn_bwd_fused_impl(Q, K, V, O, dO, LSE, Dv, dQacc, dQ, dK, dV, B, H, N, M, dh, ldq, q_bs, ldk, k_bs, ldv, v_bs, ldo,
                             o_bs, lddo, do_bs, lddq, dq_bs, lddk, dk_bs, lddv, dv_bs, scale, 1, stream_);
}
