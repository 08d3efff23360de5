// Pieces shared by the two tensor-core kernels that run a tile's layers as whole-layer GEMMs with the weights of one
// layer in shared memory (sdf_smooth_tc.cu: primal + tangent stream; sdf_tc4.cu: two tiles): the weight blob layout
// (built by surf_build_smooth_tc_weights), operand stores, the hi/lo GEMM issue.
#pragma once
#include <cuda_fp16.h>

#include "surf_internal.cuh"
#include "tc_common.cuh"

#define ST_ROWS 128
// weight buffer (one layer)
#define ST_W_BYTES 81920
#define ST_W_SMALL 16384          // lin0 forward / reverse: one 128x32 (32x128) matrix hi | lo
#define ST_W_HID_LO 32768
#define ST_W_FEAT 65536
#define ST_W_FEAT_LO 73728
#define ST_STEPS 12               // forward lin0..lin5, reverse lin5..lin0
#define ST_BLOB_BYTES (2 * ST_W_SMALL + 10 * ST_W_BYTES)
__host__ __device__ constexpr uint32_t st_step_off(int s) {
  return s == 0 ? 0u : (s < ST_STEPS - 1 ? (uint32_t)ST_W_SMALL + (uint32_t)(s - 1) * ST_W_BYTES
                                         : (uint32_t)ST_W_SMALL + 10u * ST_W_BYTES);
}
__host__ __device__ constexpr uint32_t st_step_bytes(int s) { return (s == 0 || s == ST_STEPS - 1) ? ST_W_SMALL : ST_W_BYTES; }

__device__ __forceinline__ float st_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float st_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// softplus(beta = 100): h, h' = sigmoid(100 z), h'' = 100 h' (1 - h'), from e = exp(-|100 z|) without cancellation
__device__ __forceinline__ void st_softplus(float z, float& h, float& d1, float& d2) {
  const float e = st_ex2(fabsf(z) * -144.26950408889634f);
  const float u = 1.0f + e;
  h = fmaf(st_lg2(u), 0.0069314718055994531f, fmaxf(z, 0.f));
  const float r = __fdividef(1.0f, u);
  const float er = e * r;
  d1 = z >= 0.f ? r : er;
  d2 = 100.0f * er * r;
}

// 16 values (columns k0 .. k0+15 of my row, k0 % 16 == 0) -> primal A operand in TMEM (hi | lo words)
__device__ __forceinline__ void st_store_tmem(uint32_t t_hi, uint32_t t_lo, int k0, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) tc::split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  tc::tmem_st8(t_hi + (k0 >> 1), hi);
  tc::tmem_st8(t_lo + (k0 >> 1), lo);
}
// ... -> a K-major smem operand of 128 rows (hi at base, lo at base + lo_off): two 16-byte core-matrix rows each
__device__ __forceinline__ void st_store_smem(uint8_t* base, uint32_t lo_off, int r, int k0, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) tc::split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  uint8_t* p = base + (size_t)(k0 >> 3) * 2048 + r * 16;
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  *reinterpret_cast<uint4*>(p + lo_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  *reinterpret_cast<uint4*>(p + lo_off + 2048) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}
// one element of a K-major smem operand
__device__ __forceinline__ void st_put_half(uint8_t* base, uint32_t lo_off, int r, int k, float v) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  uint8_t* p = base + (size_t)(k >> 3) * 2048 + r * 16 + (k & 7) * 2;
  *reinterpret_cast<__half*>(p) = h;
  *reinterpret_cast<__half*>(p + lo_off) = l;
}

// D (+)= A * B^T over KSTEPS k-steps of 16, fp16 hi/lo split.  TS: A in TMEM (a_hi / a_lo column addresses); else A in
// smem (a_hi / a_lo shared-window byte addresses, K-major, 128 rows).  b_addr: shared-window address of the hi matrix
// (N rows x 16 KSTEPS), lo matrix b_lo_off bytes behind it.
template <int N, int KSTEPS, bool TS>
__device__ __forceinline__ void st_gemm(uint32_t tD, uint32_t a_hi, uint32_t a_lo, uint32_t b_addr, uint32_t b_lo_off,
                                        bool acc_first, bool fast) {
  const uint32_t idesc = tc::idesc_f16(128, N, 0);
  const uint64_t db = tc::smem_desc_kmajor(0, N * 16, 128);
  const uint32_t bh = (uint32_t)(db >> 32);
  const uint32_t b0 = (uint32_t)db | (b_addr >> 4);
  const uint32_t bl = (uint32_t)db | ((b_addr + b_lo_off) >> 4);
  const uint64_t da = tc::smem_desc_kmajor(0, 2048, 128);
  const uint32_t ah = (uint32_t)(da >> 32);
  constexpr uint32_t B_KS = (N * 32) >> 4, A_KS = 4096 >> 4;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    if (TS) {
      if (ks == 0 && !acc_first) tc::mma_ts_w<false>(tD, a_hi, b0, bh, idesc);
      else tc::mma_ts_w<true>(tD, a_hi + ks * 8, b0 + ks * B_KS, bh, idesc);
      if (!fast) tc::mma_ts_w<true>(tD, a_lo + ks * 8, b0 + ks * B_KS, bh, idesc);
      if (!fast) tc::mma_ts_w<true>(tD, a_hi + ks * 8, bl + ks * B_KS, bh, idesc);
    } else {
      const uint32_t a0 = (uint32_t)da | (a_hi >> 4), a1 = (uint32_t)da | (a_lo >> 4);
      if (ks == 0 && !acc_first) tc::mma_ss_w<false>(tD, a0, ah, b0, bh, idesc);
      else tc::mma_ss_w<true>(tD, a0 + ks * A_KS, ah, b0 + ks * B_KS, bh, idesc);
      if (!fast) tc::mma_ss_w<true>(tD, a1 + ks * A_KS, ah, b0 + ks * B_KS, bh, idesc);
      if (!fast) tc::mma_ss_w<true>(tD, a0 + ks * A_KS, ah, bl + ks * B_KS, bh, idesc);
    }
  }
}

