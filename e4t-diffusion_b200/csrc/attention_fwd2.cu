// e4t_b200 — attention forward, two-query-tile ping-pong kernel (sm_100a, tcgen05 + TMEM + TMA).
// Reference call site: F.scaled_dot_product_attention in AttnProcessor2_0 (e4t/models/cross_attention.py:521-531).
//
// Why a second forward kernel.  At the level-0 shape of the SD-v1.4 UNet (N = M = 4096, 8 heads x 40) one 128x128 score
// block needs 16384 exponentials but only ~390 clk of tensor-core time: the kernel is bound by the softmax threads, not
// by the MMAs.  attn_fwd_kernel (attention.cu) covered the tensor-core/softmax hand-off latency with two CTAs per SM and
// 16 narrow softmax warps that exchange row maxima through shared memory; ncu showed 26 % of its samples spinning on
// the S-ready barrier and 17 % in idle role warps.  Here ONE CTA per SM owns NT 128-query tiles:
//
//   warp 0            TMA producer   Q_0..Q_{NT-1} once; K_j / V_j ring (kst stages of BKV keys)
//   warp 1            MMA issuer     S_t = Q_t K_j^T  and  O_t += P_t V_j  for t = 0..NT-1, interleaved so that the tensor
//                                    core works on tile t while the softmax warps of the other tiles are busy
//   warps 4+4t..7+4t  softmax tile t one query row per THREAD (TMEM lane == row): the whole BKV-wide score row sits in
//                                    registers, so there is no cross-thread max exchange, no shuffle
//
//   TMEM  S_t [t BKV, (t+1) BKV) | O_t [NT BKV + t dpad, ..)                       NT (BKV + dpad) <= 512
//         P_t (bf16, two per 32-bit column) overwrites columns [0, BKV/2) of S_t and is consumed by O_t += P_t V_j as
//         the TMEM A operand (tcgen05.mma [d], [a], bdesc): P never touches shared memory.  The tensor pipe executes in
//         issue order, so S_t(j+1) = Q_t K_{j+1}^T, issued after P_t(j) V_j, cannot overwrite P_t(j) early.
//
//   Two shapes are instantiated.  <NT 2, BKV 128> (12 warps, 168 registers): two warps per SM sub-partition; ncu (r02)
//   showed that ONE warp per sub-partition cannot keep the MUFU pipe busy in an in-order issue stream (61 % inside the
//   exponential phase: every fixed-latency dependency is exposed), and two warps that run in lock-step both wait 31 % of
//   the time for the tensor core.  <NT 4, BKV 64> (20 warps, 96 registers, head dim <= 64) keeps FOUR row-private warps
//   per sub-partition in flight at different phases of different tiles.
//
// Softmax details
//   * lazy rescale with a threshold: the running reference max m only moves when a block's max exceeds it by more than
//     2^8 (in the exp2 domain); P values may then reach 256, harmless in fp32/bf16.  After the first blocks the O
//     accumulator is practically never rescaled.
//   * exp2 split over two pipes (template POLY8 = pairs out of every 8 handled on the FMA pipe): the MUFU unit delivers
//     16 ex2/clk/SM, i.e. 1024 clk per 128x128 block, 2.6x the tensor time at dh = 40.  The FMA-pipe path is Cody-Waite
//     range reduction + a degree-3 minimax polynomial (rel. error 7.5e-5, far below bf16 rounding of P) with packed
//     fma.rn.f32x2 / add.rn.f32x2; the exponent is spliced in with one integer shift-add.
//   * row max with 3-input max (FMNMX3), scale-and-subtract with packed FFMA2.
#include "attn_common.cuh"
#include <stdlib.h>

typedef unsigned long long u64;

__device__ __forceinline__ u64 f2_pack(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 f2_add(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
      "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// 2^x for a PAIR of inputs on the FMA/ALU pipes.  x <= ~8 (threshold rescale) and may be -inf (masked keys).
//   x  = max(x, -125)                       clamp: the spliced exponent must stay >= 1
//   t  = x + 1.5*2^23                        low mantissa bits of t = round-to-nearest(x) (two's complement)
//   f  = x - (t - 1.5*2^23)  in [-0.5, 0.5]
//   2^f ~ c0 + f (c1 + f (c2 + f c3))        |rel err| < 7.5e-5
//   result = bits(2^f) + (bits(t) << 23)     adds round(x) to the exponent field
__device__ __forceinline__ u64 exp2_poly2(u64 x2) {
  float x0, x1;
  f2_unpack(x2, x0, x1);
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  const u64 xc = f2_pack(x0, x1);
  const u64 magic = f2_pack(12582912.f, 12582912.f), nmagic = f2_pack(-12582912.f, -12582912.f);
  const u64 t2 = f2_add(xc, magic);
  const u64 xi = f2_add(t2, nmagic);
  const u64 f2 = f2_fma(xi, f2_pack(-1.f, -1.f), xc);
  u64 p = f2_fma(f2_pack(0.0551714502f, 0.0551714502f), f2, f2_pack(0.242610843f, 0.242610843f));
  p = f2_fma(p, f2, f2_pack(0.693260992f, 0.693260992f));
  p = f2_fma(p, f2, f2_pack(0.999928091f, 0.999928091f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t2, t0, t1);
  const uint32_t r0 = __float_as_uint(p0) + (__float_as_uint(t0) << 23);
  const uint32_t r1 = __float_as_uint(p1) + (__float_as_uint(t1) << 23);
  return f2_pack(__uint_as_float(r0), __uint_as_float(r1));
}

// POLY8: pairs (of every 8 consecutive pairs) whose exp2 runs on the FMA pipe; 0 = all on MUFU.
template <int NT, int BKV, int POLY8>
__global__ void __launch_bounds__(128 + 128 * NT, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs a) {
  constexpr int NP = BKV / 2;                     // score pairs per row and block (= 32-bit P columns)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int q_bytes = a.DC * 16384;               // one 128-query tile
  const int kv_chunk = BKV * 128;                 // one 64-wide d chunk of a BKV-key block
  const int kv_tile = a.DC * kv_chunk;
  uint8_t* sQ = smem;                             // [NT tiles][DC][128][64]
  uint8_t* sK = sQ + NT * q_bytes;                // [kst][DC][BKV][64]
  uint8_t* sV = sK + a.kst * kv_tile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + a.kst * kv_tile);
  uint64_t* q_full = bars;            // [1]
  uint64_t* k_full = bars + 1;        // [4]
  uint64_t* k_empty = k_full + 4;     // [4]
  uint64_t* v_full = k_empty + 4;     // [4]
  uint64_t* v_empty = v_full + 4;     // [4]
  uint64_t* s_full = v_empty + 4;     // [NT]  MMA -> softmax t : S_t(j) complete (and everything issued before it)
  uint64_t* p_ready = s_full + NT;    // [NT]  softmax t -> MMA : P_t(j) written, O_t rescaled
  uint64_t* o_done = p_ready + NT;    // [NT]  MMA -> softmax t : last P_t V retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + NT);
  float* s_zero = reinterpret_cast<float*>(tmem_slot + 2);   // 0.0f: read after the exp-phase token (ordering anchor)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (128 * NT), h = blockIdx.y, b = blockIdx.z;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < NT; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_done[i], 1);
    }
    *s_zero = 0.f;
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tO0 = tmem + (uint32_t)(NT * BKV);

  // Both role warps run their loops in ONE elected thread with descriptors built once and the ring stage / phase tracked
  // incrementally (see attn_fwd3_kernel: the issue rate of the MMA thread is on the critical path).
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, (uint32_t)(NT * q_bytes));
      for (int t = 0; t < NT; ++t)
        for (int c = 0; c < a.DC; ++c) tma_load_4d(sQ + t * q_bytes + c * 16384, &mapQ, q_full, c * 64, h, q0 + 128 * t, b);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < a.nblk; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1u);
        mbar_expect_tx(&k_full[st], (uint32_t)kv_tile);
        for (int c = 0; c < a.DC; ++c)
          tma_load_4d(sK + st * kv_tile + c * kv_chunk, &mapK, &k_full[st], c * 64, h, j * BKV, b);
        mbar_wait(&v_empty[st], ph ^ 1u);
        mbar_expect_tx(&v_full[st], (uint32_t)kv_tile);
        for (int c = 0; c < a.DC; ++c)
          tma_load_4d(sV + st * kv_tile + c * kv_chunk, &mapV, &v_full[st], c * 64, h, j * BKV, b);
        if (++st == a.kst) {
          st = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16((uint32_t)BKV, false, false);
      const uint32_t idesc_o = umma_idesc_bf16((uint32_t)a.dpad, false, true);
      // descriptor bases in 16-byte units: Q tile t: + t * (q_bytes >> 4); k-step kk of dh: chunk (kk >> 2), 32 B * (kk & 3)
      const uint64_t dQ0 = umma_desc(smem_u32(sQ), 16, 1024);
      const uint64_t dK0 = umma_desc(smem_u32(sK), 16, 1024);
      const uint64_t dV0 = umma_desc(smem_u32(sV), (uint32_t)kv_chunk, 1024);
      const uint32_t q_units = (uint32_t)q_bytes >> 4, stage_u = (uint32_t)kv_tile >> 4, kchunk_u = (uint32_t)kv_chunk >> 4;
      const int ksteps = (a.dh + 15) >> 4;
      auto issue_qk = [&](uint32_t tS, int t, uint64_t dK) {
        for (int kk = 0; kk < ksteps; ++kk)
          umma_bf16(tS, dQ0 + (uint64_t)((uint32_t)t * q_units + (uint32_t)(kk >> 2) * 1024u + (uint32_t)(kk & 3) * 2u),
                    dK + (uint64_t)((uint32_t)(kk >> 2) * kchunk_u + (uint32_t)(kk & 3) * 2u), idesc_s, kk > 0 ? 1u : 0u);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < NT; ++t) {
        issue_qk(tmem + (uint32_t)(t * BKV), t, dK0);
        if (t == NT - 1) umma_commit(&k_empty[0]);
        umma_commit(&s_full[t]);
      }
      int st = 0, stn = a.kst > 1 ? 1 : 0;
      uint32_t ph = 0, phn = a.kst > 1 ? 0u : 1u;     // stage / phase of block j and of block j + 1
      for (int j = 0; j < a.nblk; ++j) {
        const bool more = j + 1 < a.nblk;
        const uint64_t dV = dV0 + (uint64_t)((uint32_t)st * stage_u);
        const uint64_t dK = dK0 + (uint64_t)((uint32_t)stn * stage_u);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          mbar_wait(&p_ready[t], (uint32_t)(j & 1));
          if (t == 0) {
            mbar_wait(&v_full[st], ph);
            if (more) mbar_wait(&k_full[stn], phn);
          }
          tc_fence_after();
          const uint32_t tS = tmem + (uint32_t)(t * BKV);
          const uint32_t tO = tO0 + (uint32_t)(t * a.dpad);
#pragma unroll
          for (int ks = 0; ks < BKV / 16; ++ks)
            umma_bf16_ts(tO, tS + (uint32_t)(ks * 8), dV + (uint64_t)(ks * 128), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          if (t == NT - 1) umma_commit(&v_empty[st]);
          if (more) {
            issue_qk(tS, t, dK);
            if (t == NT - 1) umma_commit(&k_empty[stn]);
            umma_commit(&s_full[t]);
          } else {
            umma_commit(&o_done[t]);
          }
        }
        st = stn;
        ph = phn;
        if (++stn == a.kst) {
          stn = 0;
          phn ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax warps: tile t = (warp - 4) / 4, one query row per thread =====================
    const int t = (warp - 4) >> 2;
    const int ew = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row = ew * 32 + lane;
    const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
    const uint32_t tS = tmem + (uint32_t)(t * BKV) + lane_base;
    const uint32_t tO = tO0 + (uint32_t)(t * a.dpad) + lane_base;
    const float sl2 = a.scale * kLog2e;
    const u64 sl2_2 = f2_pack(sl2, sl2);
    const int ochunk = a.dpad >> 4;
    const bool token = (NT == 2) && a.pbuf;
    float m = -INFINITY;                           // reference max of this row (raw score units)
    float l0 = 0.f, l1 = 0.f;
    if (token && t == 1) asm volatile("bar.arrive 2, 256;" ::: "memory");   // tile 0 owns the first exp phase
    for (int j = 0; j < a.nblk; ++j) {
      mbar_wait(&s_full[t], (uint32_t)(j & 1));
      tc_fence_after();
      uint32_t v[BKV];
#pragma unroll
      for (int c = 0; c < BKV; c += 32) tmem_ld32(tS + (uint32_t)c, v + c);
      tmem_ld_wait();
      const int kv0 = j * BKV;
      if (kv0 + BKV > a.M) {                       // ragged last block: keys >= M are masked out
#pragma unroll
        for (int e = 0; e < BKV; ++e)
          if (kv0 + e >= a.M) v[e] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int e = 0; e < BKV; e += 4) {
        mx0 = max3(mx0, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
        mx1 = max3(mx1, __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
      }
      const float mx = fmaxf(mx0, mx1);
      // move the reference max only when this block exceeds it by more than 2^8 in the exp2 domain
      const bool need = (mx - m) * sl2 > 8.f;      // m = -inf on the first block -> true
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mx : m;
        const float alpha = ex2_approx((m - m_new) * sl2);   // 1 for rows that keep their max; 0 on the first block
        l0 *= alpha;
        l1 *= alpha;
        m = m_new;
        if (j > 0) {   // O_t holds P V of blocks < j (complete: s_full(j) was committed after P_t(j-1) V)
          for (int oc = 0; oc < ochunk; ++oc) {
            uint32_t ov[16];
            tmem_ld16(tO + (uint32_t)(oc * 16), ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
            tmem_st16(tO + (uint32_t)(oc * 16), ov);
          }
        }
      }
      float nmb = -m * sl2;
      // exp-phase token (NT == 2 only): the two softmax warpgroups take turns on the MUFU pipe.  ptxas is free to move
      // pure register arithmetic across a bar.sync, so the exponential phase is made to depend on a (zero) value
      // loaded from shared memory AFTER the barrier.
      if (token) {
        if (t == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
        else asm volatile("bar.sync 3, 256;" ::: "memory");
        float z;
        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(z) : "r"(smem_u32(s_zero)) : "memory");
        nmb += z;
      }
      const u64 nmb2 = f2_pack(nmb, nmb);
      u64 ls = f2_pack(0.f, 0.f), ls2 = f2_pack(0.f, 0.f);
      // Software-pipelined by hand (the SM issues in order): the MUFU results of pair e are consumed (row sum, bf16
      // pack) kLag pairs later, and the scale-and-subtract FFMA2 of pair e + kLead is issued in between.
      constexpr int kLag = 6, kLead = 4;
      if constexpr (POLY8 != 0) {
#pragma unroll
        for (int e = 0; e < NP; ++e) {               // pair e = columns 2e, 2e+1
          const u64 x2 = f2_fma(f2_pack(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), sl2_2, nmb2);
          u64 p2;
          if ((e & 7) < POLY8) {
            p2 = exp2_poly2(x2);
          } else {
            float x0, x1, p0, p1;
            f2_unpack(x2, x0, x1);
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
            p2 = f2_pack(p0, p1);
          }
          float p0, p1;
          f2_unpack(p2, p0, p1);
          v[e] = pack_bf16(p0, p1);
          if (e & 1) ls2 = f2_add(ls2, p2);
          else ls = f2_add(ls, p2);
        }
      } else {
        u64 xq[NP];                                  // x of pair e, then p of pair e (registers of v[] are recycled)
#pragma unroll
        for (int e = 0; e < kLead; ++e)
          xq[e] = f2_fma(f2_pack(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), sl2_2, nmb2);
#pragma unroll
        for (int e = 0; e < NP + kLag; ++e) {
          if (e < NP) {
            float x0, x1, p0, p1;
            f2_unpack(xq[e], x0, x1);
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
            if (e + kLead < NP)
              xq[e + kLead] = f2_fma(f2_pack(__uint_as_float(v[2 * (e + kLead)]), __uint_as_float(v[2 * (e + kLead) + 1])),
                                     sl2_2, nmb2);
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
            xq[e] = f2_pack(p0, p1);
          }
          if (e >= kLag) {
            const int c = e - kLag;
            if (c & 1) ls2 = f2_add(ls2, xq[c]);
            else ls = f2_add(ls, xq[c]);
            float p0, p1;
            f2_unpack(xq[c], p0, p1);
            v[c] = pack_bf16(p0, p1);
          }
        }
      }
      if (token && !(t == 1 && j == a.nblk - 1)) {   // hand the token to the other warpgroup
        if (t == 0) asm volatile("bar.arrive 3, 256;" ::: "memory");
        else asm volatile("bar.arrive 2, 256;" ::: "memory");
      }
      {
        float s0, s1;
        f2_unpack(f2_add(ls, ls2), s0, s1);
        l0 += s0;
        l1 += s1;
      }
      // P_t -> TMEM columns [0, BKV/2) of S_t (A operand of O_t += P_t V)
#pragma unroll
      for (int c = 0; c < NP; c += 32) tmem_st32(tS + (uint32_t)c, v + c);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[t]);
    }
    // ---- epilogue: normalise and store this tile's rows ----
    mbar_wait(&o_done[t], 0);
    tc_fence_after();
    const float lt = l0 + l1;
    const float inv_l = 1.f / lt;
    const int n = q0 + 128 * t + row;
    for (int oc = 0; oc < ochunk; ++oc) {
      const int c = oc * 16;
      uint32_t ov[16];
      tmem_ld16(tO + (uint32_t)c, ov);
      tmem_ld_wait();
      if (n < a.N) {
        bf16* o = a.O + (long long)b * a.o_bs + (long long)n * a.ldo + h * a.dh + c;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          if (c + i < a.dh) {
            *reinterpret_cast<uint4*>(o + i) =
                make_uint4(pack_bf16(__uint_as_float(ov[i]) * inv_l, __uint_as_float(ov[i + 1]) * inv_l),
                           pack_bf16(__uint_as_float(ov[i + 2]) * inv_l, __uint_as_float(ov[i + 3]) * inv_l),
                           pack_bf16(__uint_as_float(ov[i + 4]) * inv_l, __uint_as_float(ov[i + 5]) * inv_l),
                           pack_bf16(__uint_as_float(ov[i + 6]) * inv_l, __uint_as_float(ov[i + 7]) * inv_l));
          }
        }
      }
    }
    if (n < a.N) a.LSE[((long long)b * a.H + h) * a.N + n] = m * a.scale + logf(lt);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================
// Double-buffered S (attn_fwd3_kernel): two 128-query tiles, each with TWO score buffers in TMEM.
//   TMEM  S_t,b [ (2t+b) BKV, .. ) | O_t [4 BKV + t dpad, ..)            2 (2 BKV + dpad) <= 512  ->  BKV = 96 for dpad <= 64
// The tensor core computes S_t(j+2) right after P_t(j)·V_j, i.e. TWO blocks ahead of the softmax, so the softmax warps
// of a tile run block after block without ever waiting for S (in attn_fwd2_kernel S_t(j+1) can only be issued after
// P_t(j) exists: the softmax warps waited 31 % of the time for it, and the MUFU pipe idled 45 %).  The only remaining
// softmax -> tensor -> softmax dependency is the (rare) rescale of O, which waits for P_t(j-1)·V on its own barrier.
// P_t(j) overwrites the first BKV/2 columns of S_t,(j&1); S_t(j+2) is issued after P_t(j)·V_j and the tensor pipe
// executes in order, so it cannot overwrite P_t(j) early.
// =============================================================================================
template <int BKV, int POLY8>
__global__ void __launch_bounds__(384, 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs a) {
  constexpr int NP = BKV / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int q_bytes = a.DC * 16384;
  const int kv_chunk = BKV * 128;
  const int kv_tile = a.DC * kv_chunk;
  uint8_t* sQ = smem;                             // [2 tiles][DC][128][64]
  uint8_t* sK = sQ + 2 * q_bytes;                 // [kst][DC][BKV][64]
  uint8_t* sV = sK + a.kst * kv_tile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + a.kst * kv_tile);
  uint64_t* q_full = bars;            // [1]
  uint64_t* k_full = bars + 1;        // [4]
  uint64_t* k_empty = k_full + 4;     // [4]
  uint64_t* v_full = k_empty + 4;     // [4]
  uint64_t* v_empty = v_full + 4;     // [4]
  uint64_t* s_full = v_empty + 4;     // [2 tiles][2 buffers]  MMA -> softmax t : S_t(j) complete in buffer j & 1
  uint64_t* p_ready = s_full + 4;     // [2 tiles][2 buffers]  softmax t -> MMA : P_t(j) written (and O_t rescaled).  One
                                      // barrier per S buffer: a softmax warp group can be TWO blocks ahead of the MMA warp
                                      // (S_t(j+1) exists while the MMA warp still waits for the other tile's P(j-1)), and a
                                      // parity wait only tolerates a lead of one phase (a single barrier deadlocked, r02 call 10)
  uint64_t* pv_done = p_ready + 4;    // [2]  MMA -> softmax t : P_t(j) V retired (only consulted before a rescale / at the end)
  uint64_t* o_done = pv_done + 2;     // [2]  MMA -> softmax t : the LAST P_t V retired.  Not pv_done: after its last block a
                                      // softmax group may be two P·V behind the tensor core, three possible phase counts alias
                                      // under a parity wait (the epilogue read O early: right LSE, wrong O; r02 call 12)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tO0 = tmem + (uint32_t)(4 * BKV);

  // The two role warps run their whole loop in ONE elected thread, with a compile-time ring depth and descriptors
  // precomputed outside the loop: the first version (whole-warp control flow, `j % a.kst`, descriptors rebuilt per k-step)
  // executed ~700 instructions per block pair in the MMA warp, and ncu showed that warp BUSY (not waiting) while the
  // softmax warps waited a third of their time for S — the issue rate of this one thread is on the critical path.
  constexpr int KST = 4;
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, (uint32_t)(2 * q_bytes));
      for (int t = 0; t < 2; ++t) tma_load_4d(sQ + t * q_bytes, &mapQ, q_full, 0, h, q0 + 128 * t, b);
      // K runs two blocks ahead of V (S is computed two blocks ahead): K(0), K(1), then K(j+2), V(j) per block
      const int nsteps = a.nblk + 2;
      for (int step = 0; step < nsteps; ++step) {
        if (step < a.nblk) {
          const int st = step & (KST - 1);
          mbar_wait(&k_empty[st], (uint32_t)(((step >> 2) & 1) ^ 1));
          mbar_expect_tx(&k_full[st], (uint32_t)kv_tile);
          tma_load_4d(sK + st * kv_tile, &mapK, &k_full[st], 0, h, step * BKV, b);
        }
        const int jv = step - 2;
        if (jv >= 0) {
          const int st = jv & (KST - 1);
          mbar_wait(&v_empty[st], (uint32_t)(((jv >> 2) & 1) ^ 1));
          mbar_expect_tx(&v_full[st], (uint32_t)kv_tile);
          tma_load_4d(sV + st * kv_tile, &mapV, &v_full[st], 0, h, jv * BKV, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16((uint32_t)BKV, false, false);
      const uint32_t idesc_o = umma_idesc_bf16((uint32_t)a.dpad, false, true);
      const int ksteps = (a.dh + 15) >> 4;                              // dh <= 64: one 64-wide chunk
      const uint64_t dQ0 = umma_desc(smem_u32(sQ), 16, 1024);           // tile t: + t * (16384 >> 4); k-step k: + 2 k
      const uint64_t dK0 = umma_desc(smem_u32(sK), 16, 1024);           // stage s: + s * stage_u
      const uint64_t dV0 = umma_desc(smem_u32(sV), (uint32_t)kv_chunk, 1024);   // 16-key step ks: + ks * (2048 >> 4)
      const uint32_t stage_u = (uint32_t)kv_tile >> 4;
      mbar_wait(q_full, 0);
      // prologue: S_t(0) and S_t(1)
      for (int jj = 0; jj < 2 && jj < a.nblk; ++jj) {
        mbar_wait(&k_full[jj], 0);
        tc_fence_after();
        for (int t = 0; t < 2; ++t) {
          const uint32_t tS = tmem + (uint32_t)((2 * t + jj) * BKV);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tS, dQ0 + (uint64_t)(t * 1024 + 2 * k), dK0 + (uint64_t)(jj * stage_u + 2 * k), idesc_s, k > 0);
          if (t == 1) umma_commit(&k_empty[jj]);
          umma_commit(&s_full[2 * t + jj]);
        }
      }
      for (int j = 0; j < a.nblk; ++j) {
        const int st = j & (KST - 1), b01 = j & 1;
        const int jn = j + 2, stn = jn & (KST - 1);
        const bool more = jn < a.nblk;
        const uint32_t ph2 = (uint32_t)((j >> 1) & 1);
        const uint64_t dV = dV0 + (uint64_t)((uint32_t)st * stage_u);
        const uint64_t dK = dK0 + (uint64_t)((uint32_t)stn * stage_u);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_ready[2 * t + b01], ph2);
          if (t == 0) {
            mbar_wait(&v_full[st], (uint32_t)((j >> 2) & 1));
            if (more) mbar_wait(&k_full[stn], (uint32_t)((jn >> 2) & 1));
          }
          tc_fence_after();
          const uint32_t tS = tmem + (uint32_t)((2 * t + b01) * BKV);
          const uint32_t tO = tO0 + (uint32_t)(t * a.dpad);
#pragma unroll
          for (int ks = 0; ks < BKV / 16; ++ks)
            umma_bf16_ts(tO, tS + (uint32_t)(ks * 8), dV + (uint64_t)(ks * 128), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          if (t == 1) umma_commit(&v_empty[st]);
          umma_commit(&pv_done[t]);
          if (j == a.nblk - 1) umma_commit(&o_done[t]);
          if (more) {
            for (int k = 0; k < ksteps; ++k)
              umma_bf16(tS, dQ0 + (uint64_t)(t * 1024 + 2 * k), dK + (uint64_t)(2 * k), idesc_s, k > 0);
            if (t == 1) umma_commit(&k_empty[stn]);
            umma_commit(&s_full[2 * t + b01]);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax warps: tile t = (warp - 4) / 4, one query row per thread =====================
    const int t = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
    const uint32_t tO = tO0 + (uint32_t)(t * a.dpad) + lane_base;
    const float sl2 = a.scale * kLog2e;
    const u64 sl2_2 = f2_pack(sl2, sl2);
    const int ochunk = a.dpad >> 4;
    float m = -INFINITY;
    float l0 = 0.f, l1 = 0.f;
    for (int j = 0; j < a.nblk; ++j) {
      const uint32_t tS = tmem + (uint32_t)((2 * t + (j & 1)) * BKV) + lane_base;
      mbar_wait(&s_full[2 * t + (j & 1)], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      uint32_t v[BKV];
#pragma unroll
      for (int c = 0; c < BKV; c += 32) tmem_ld32(tS + (uint32_t)c, v + c);
      tmem_ld_wait();
      const int kv0 = j * BKV;
      if (kv0 + BKV > a.M) {
#pragma unroll
        for (int e = 0; e < BKV; ++e)
          if (kv0 + e >= a.M) v[e] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int e = 0; e < BKV; e += 4) {
        mx0 = max3(mx0, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
        mx1 = max3(mx1, __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
      }
      const float mx = fmaxf(mx0, mx1);
      const bool need = (mx - m) * sl2 > 8.f;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mx : m;
        const float alpha = ex2_approx((m - m_new) * sl2);
        l0 *= alpha;
        l1 *= alpha;
        m = m_new;
        if (j > 0) {
          mbar_wait(&pv_done[t], (uint32_t)((j - 1) & 1));   // O_t must hold P V of every block < j
          tc_fence_after();
          for (int oc = 0; oc < ochunk; ++oc) {
            uint32_t ov[16];
            tmem_ld16(tO + (uint32_t)(oc * 16), ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
            tmem_st16(tO + (uint32_t)(oc * 16), ov);
          }
        }
      }
      const float nmb = -m * sl2;
      const u64 nmb2 = f2_pack(nmb, nmb);
      u64 ls = f2_pack(0.f, 0.f), ls2 = f2_pack(0.f, 0.f);
      constexpr int kLag = 6, kLead = 4;
      if constexpr (POLY8 != 0) {
        // exp2 of POLY8 pairs out of every 8 on the FMA pipe (exp2_poly2): with S double-buffered the MUFU pipe is ~70 %
        // busy, so moving a share of the exponentials off it shortens the phase (it did not while the kernel waited for S)
#pragma unroll
        for (int e = 0; e < NP; ++e) {
          const u64 x2 = f2_fma(f2_pack(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), sl2_2, nmb2);
          u64 p2;
          if ((e & 7) < POLY8) {
            p2 = exp2_poly2(x2);
          } else {
            float x0, x1, p0, p1;
            f2_unpack(x2, x0, x1);
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
            p2 = f2_pack(p0, p1);
          }
          float p0, p1;
          f2_unpack(p2, p0, p1);
          v[e] = pack_bf16(p0, p1);
          if (e & 1) ls2 = f2_add(ls2, p2);
          else ls = f2_add(ls, p2);
        }
      } else {
        u64 xq[NP];
#pragma unroll
        for (int e = 0; e < kLead; ++e)
          xq[e] = f2_fma(f2_pack(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), sl2_2, nmb2);
#pragma unroll
        for (int e = 0; e < NP + kLag; ++e) {
          if (e < NP) {
            float x0, x1, p0, p1;
            f2_unpack(xq[e], x0, x1);
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
            if (e + kLead < NP)
              xq[e + kLead] = f2_fma(f2_pack(__uint_as_float(v[2 * (e + kLead)]), __uint_as_float(v[2 * (e + kLead) + 1])),
                                     sl2_2, nmb2);
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
            xq[e] = f2_pack(p0, p1);
          }
          if (e >= kLag) {
            const int c = e - kLag;
            if (c & 1) ls2 = f2_add(ls2, xq[c]);
            else ls = f2_add(ls, xq[c]);
            float p0, p1;
            f2_unpack(xq[c], p0, p1);
            v[c] = pack_bf16(p0, p1);
          }
        }
      }
      {
        float s0, s1;
        f2_unpack(f2_add(ls, ls2), s0, s1);
        l0 += s0;
        l1 += s1;
      }
      // P_t(j) -> TMEM columns [0, BKV/2) of S_t,(j&1)
      tmem_st32(tS, v);
      if constexpr (NP > 32) tmem_st16(tS + 32u, v + 32);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[2 * t + (j & 1)]);
    }
    // ---- epilogue ----
    mbar_wait(&o_done[t], 0);
    tc_fence_after();
    const float lt = l0 + l1;
    const float inv_l = 1.f / lt;
    const int n = q0 + 128 * t + row;
    for (int oc = 0; oc < ochunk; ++oc) {
      const int c = oc * 16;
      uint32_t ov[16];
      tmem_ld16(tO + (uint32_t)c, ov);
      tmem_ld_wait();
      if (n < a.N) {
        bf16* o = a.O + (long long)b * a.o_bs + (long long)n * a.ldo + h * a.dh + c;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          if (c + i < a.dh) {
            *reinterpret_cast<uint4*>(o + i) =
                make_uint4(pack_bf16(__uint_as_float(ov[i]) * inv_l, __uint_as_float(ov[i + 1]) * inv_l),
                           pack_bf16(__uint_as_float(ov[i + 2]) * inv_l, __uint_as_float(ov[i + 3]) * inv_l),
                           pack_bf16(__uint_as_float(ov[i + 4]) * inv_l, __uint_as_float(ov[i + 5]) * inv_l),
                           pack_bf16(__uint_as_float(ov[i + 6]) * inv_l, __uint_as_float(ov[i + 7]) * inv_l));
          }
        }
      }
    }
    if (n < a.N) a.LSE[((long long)b * a.H + h) * a.N + n] = m * a.scale + logf(lt);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Host
// ---------------------------------------------------------------------------------------------
int e4t_attn_make_head_map(CUtensorMap* m, const void* p, int dh, int H, int rows, int B, long long ld, long long bs,
                           int box_rows);

template <int NT, int BKV, int POLY8>
static int launch_fwd2(const CUtensorMap& mQ, const CUtensorMap& mK, const CUtensorMap& mV, const AttnArgs& a, size_t smem,
                       cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_fwd2_kernel<NT, BKV, POLY8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return -1;
    attr = true;
  }
  const dim3 grid(cdiv(a.N, 128 * NT), a.H, a.B);
  attn_fwd2_kernel<NT, BKV, POLY8><<<grid, 128 + 128 * NT, smem, st>>>(mQ, mK, mV, a);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Returns 1 if this kernel took the call, 0 if the shape is outside its envelope (caller falls back), <0 on error.
int e4t_attn_fwd2_try(const void* Q, const void* K, const void* V, void* O, float* LSE, int B, int H, int N, int M, int dh,
                      long long ldq, long long q_bs, long long ldk, long long k_bs, long long ldv, long long v_bs,
                      long long ldo, long long o_bs, float scale, cudaStream_t st) {
  // E4T_ATTN_FWD2: "0" disables these kernels (single-tile attn_fwd_kernel); otherwise a string of flags:
  //   'd' / 's'  force the double-buffered-S kernel attn_fwd3_kernel (default where it applies: dh <= 64, M >= 192) /
  //              the single-buffered two-tile kernel attn_fwd2_kernel
  //   'p<k>'     FMA-pipe exp2 for k of every 8 pairs (default: 2 in attn_fwd3_kernel — 0.701 vs 0.723 ms at level 0, r02
  //              call 15 — and 0 in attn_fwd2_kernel, which waits for S rather than for the MUFU pipe)
  //   'f'        fwd2 only: the <4 tiles x 64 keys> shape (measured slower, see below);  'n'  fwd2 only: no exp-phase token
  const char* e = getenv("E4T_ATTN_FWD2");
  int poly8 = -1, wide = 1, token = 1, dbuf = 1;
  if (e) {
    if (e[0] == '0' && e[1] == 0) return 0;
    for (const char* c = e; *c; ++c) {
      if (*c == 'p' && c[1] >= '0' && c[1] <= '7') poly8 = c[1] - '0';
      if (*c == 'f') wide = 0;
      if (*c == 'n') token = 0;
      if (*c == 'd') dbuf = 1;
      if (*c == 's') dbuf = 0;
    }
  }
  if (dh > 128 || M < 128 || N < 128) return 0;
  if (dbuf && dh <= 64 && M >= 192) {   // attn_fwd3_kernel<96>: two S buffers per tile, 96-key blocks
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.H = H; a.N = N; a.M = M; a.dh = dh;
    a.DC = 1;
    a.dpad = (dh + 15) / 16 * 16;
    a.BKV = 96;
    a.nblk = cdiv(M, 96);
    a.kst = 4;
    a.scale = scale;
    a.O = (bf16*)O; a.ldo = ldo; a.o_bs = o_bs; a.LSE = LSE;
    const size_t smem = (size_t)2 * 16384 + (size_t)2 * 4 * 96 * 128 + 512 + 1024;
    CUtensorMap mQ, mK, mV;
    if (e4t_attn_make_head_map(&mQ, Q, dh, H, N, B, ldq, q_bs, 128)) return -1;
    if (e4t_attn_make_head_map(&mK, K, dh, H, M, B, ldk, k_bs, 96)) return -1;
    if (e4t_attn_make_head_map(&mV, V, dh, H, M, B, ldv, v_bs, 96)) return -1;
    const dim3 grid3(cdiv(N, 256), H, B);
    static bool attr3 = false;
    if (!attr3) {
      if (cudaFuncSetAttribute(attn_fwd3_kernel<96, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(attn_fwd3_kernel<96, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(attn_fwd3_kernel<96, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(attn_fwd3_kernel<96, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return -1;
      attr3 = true;
    }
    switch (poly8 < 0 ? 2 : poly8) {
      case 1: attn_fwd3_kernel<96, 1><<<grid3, 384, smem, st>>>(mQ, mK, mV, a); break;
      case 2: attn_fwd3_kernel<96, 2><<<grid3, 384, smem, st>>>(mQ, mK, mV, a); break;
      case 3: attn_fwd3_kernel<96, 3><<<grid3, 384, smem, st>>>(mQ, mK, mV, a); break;
      default: attn_fwd3_kernel<96, 0><<<grid3, 384, smem, st>>>(mQ, mK, mV, a); break;
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.N = N; a.M = M; a.dh = dh;
  a.DC = cdiv(dh, 64);
  a.dpad = (dh + 15) / 16 * 16;
  // default: two 128-query tiles x 128-key blocks.  The four-tile x 64-key shape (flag 'f'; accumulators fit when
  // 4 x (64 + dpad) <= 512) was built to put four row-private softmax warps on every sub-partition, but measured
  // SLOWER at the level-0 shape (0.947 ms vs 0.906 ms, r02 call 6): with 64-key blocks the per-tile dependency loop
  // softmax -> P ready -> P·V, Q·K^T issue -> S ready is paid twice as often and its latency (mbarrier wake-ups, ~600 clk
  // tcgen05.ld under MMA load), not MUFU throughput, then dominates — the softmax warps wait 67 % of the time for S.
  const bool four = !wide && a.dpad <= 64 && N >= 384;
  const int NT = four ? 4 : 2, BKV = four ? 64 : 128;
  a.BKV = BKV;
  a.nblk = cdiv(M, BKV);
  a.scale = scale;
  a.pbuf = token;     // (field reused) 1 = exp-phase token ping-pong between the two softmax warpgroups (NT == 2)
  a.O = (bf16*)O; a.ldo = ldo; a.o_bs = o_bs; a.LSE = LSE;
  const size_t fixed = (size_t)NT * a.DC * 16384 + 512 + 1024;
  const size_t per_stage = (size_t)2 * a.DC * BKV * 128;
  int kst = (int)((227 * 1024 - fixed) / per_stage);
  if (kst > 4) kst = 4;
  if (kst > a.nblk) kst = a.nblk;
  if (kst < 1) return 0;
  a.kst = kst;
  const size_t smem = fixed + (size_t)kst * per_stage;
  CUtensorMap mQ, mK, mV;
  if (e4t_attn_make_head_map(&mQ, Q, dh, H, N, B, ldq, q_bs, 128)) return -1;
  if (e4t_attn_make_head_map(&mK, K, dh, H, M, B, ldk, k_bs, BKV)) return -1;
  if (e4t_attn_make_head_map(&mV, V, dh, H, M, B, ldv, v_bs, BKV)) return -1;
  if (four) {
    switch (poly8) {
      case 2: return launch_fwd2<4, 64, 2>(mQ, mK, mV, a, smem, st);
      case 3: return launch_fwd2<4, 64, 3>(mQ, mK, mV, a, smem, st);
      case 4: return launch_fwd2<4, 64, 4>(mQ, mK, mV, a, smem, st);
      default: return launch_fwd2<4, 64, 0>(mQ, mK, mV, a, smem, st);
    }
  }
  switch (poly8) {
    case 2: return launch_fwd2<2, 128, 2>(mQ, mK, mV, a, smem, st);
    case 4: return launch_fwd2<2, 128, 4>(mQ, mK, mV, a, smem, st);
    default: return launch_fwd2<2, 128, 0>(mQ, mK, mV, a, smem, st);
  }
}
