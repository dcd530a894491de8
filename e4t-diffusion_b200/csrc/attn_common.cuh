// e4t_b200 — pieces shared by the attention kernels (attention.cu, attention_fwd2.cu).
#pragma once
#include "common.cuh"

static constexpr float kLog2e = 1.4426950408889634f;

struct AttnArgs {
  int B, H, N, M, dh;
  int DC;        // 64-wide chunks of dh
  int dpad;      // dh rounded up to 16 (MMA N of the output GEMMs)
  int BKV;       // key/value block (fwd, dQ) or query block (dKV), multiple of 16
  int nblk;      // number of blocks looped over
  int kst;       // ring stages
  int sbuf;      // fwd: S accumulator buffers in TMEM (2 = software pipelined, 1 = rely on 2 CTAs/SM)
  int pbuf;      // fwd: P buffers in smem (2 = softmax never waits for the previous P·V to retire)
  int tmem_cols; // TMEM columns to allocate (256 lets two CTAs share an SM)
  int causal;    // fused backward only: keys after the query are masked (CLIP text tower)
  float scale;   // dh^-0.5
  // pointers / strides (elements)
  bf16* O;  long long ldo, o_bs;
  float* LSE;    // [B][H][N]
  const float* Dv;  // [B][H][N] rowsum(dO∘O)
  bf16* dQ; long long lddq, dq_bs;
  bf16* dK; long long lddk, dk_bs;
  bf16* dV; long long lddv, dv_bs;
};

__device__ __forceinline__ void mma_kmajor(uint32_t d_tmem, uint32_t sA, uint32_t a_chunk, uint32_t sB,
                                           uint32_t b_chunk, int dh, int DC, uint32_t idesc) {
  // D = A[128][dh] · B[N][dh]ᵀ, both K-major, dh split in 64-wide chunks
  uint32_t acc = 0;
  for (int c = 0; c < DC; ++c) {
    const int rem = dh - 64 * c;
    const int ks = rem >= 64 ? 4 : (rem + 15) / 16;
    for (int k = 0; k < ks; ++k) {
      umma_bf16(d_tmem, umma_desc(sA + c * a_chunk + k * 32, 16, 1024), umma_desc(sB + c * b_chunk + k * 32, 16, 1024),
                idesc, acc);
      acc = 1;
    }
  }
}
__device__ __forceinline__ void mma_pv(uint32_t d_tmem, uint32_t sP, uint32_t sV, uint32_t v_chunk, int kdim,
                                       uint32_t idesc, uint32_t acc) {
  // D[128][dpad] (+)= P[128][kdim] (K-major, 64-col chunks of 16 KiB) · V[kdim][dpad] (MN-major, 64-wide d chunks)
  for (int ks = 0; ks < kdim / 16; ++ks) {
    umma_bf16(d_tmem, umma_desc(sP + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
              umma_desc(sV + ks * 2048, v_chunk, 1024), idesc, acc);
    acc = 1;
  }
}

__device__ __forceinline__ void mma_pv_ts(uint32_t d_tmem, uint32_t tP, uint32_t sV, uint32_t v_chunk, int kdim,
                                          uint32_t idesc, uint32_t acc) {
  // as mma_pv with the A operand P[128][kdim] in TMEM: K-step ks = 8 columns (two bf16 per column)
  for (int ks = 0; ks < kdim / 16; ++ks) {
    umma_bf16_ts(d_tmem, tP + (uint32_t)(ks * 8), umma_desc(sV + ks * 2048, v_chunk, 1024), idesc, acc);
    acc = 1;
  }
}

__device__ __forceinline__ float max32(const uint32_t* v, float mx) {
#pragma unroll
  for (int e = 0; e < 32; e += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[e]), __uint_as_float(v[e + 1])));
  return mx;
}
// store 32 bf16 (packed in w[16]) of row `rowoff/128` at columns [col0, col0+32) of a K-major SWIZZLE_128B tile set
__device__ __forceinline__ void sts_row32(uint8_t* tile, uint32_t rowoff, uint32_t r7, int col0, const uint32_t* w) {
  uint8_t* pc = tile + (col0 >> 6) * 16384 + rowoff;
  const uint32_t cb = (uint32_t)((col0 & 63) >> 3);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(pc + (((cb + q) ^ r7) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

