// Self-test of the tcgen05 building blocks used by the tensor-core MLP kernels: TMEM alloc, tcgen05.st of
// an fp16 A operand (row = TMEM lane, two halves per 32-bit column), B operand in shared memory in the
// K-major no-swizzle canonical layout, tcgen05.mma (TS form), tcgen05.commit -> mbarrier, tcgen05.ld.
//   D (128 x N, fp32) = A (128 x K) * B (N x K)^T      split = 1: fp16 hi/lo split, 3 MMAs per K step
#include "surf_internal.cuh"
#include "tc_common.cuh"

__global__ void __launch_bounds__(128, 1)
k_tc_selftest(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int N, int split) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // smem: B_hi then B_lo, each (K/8) chunks x N rows x 16 B ; element (n,k) at (k/8)*N*16 + n*16 + (k%8)*2
  const uint32_t lbo = (uint32_t)N * 16, sbo = 128;
  __half* Bhi = reinterpret_cast<__half*>(smem);
  __half* Blo = Bhi + (size_t)N * K;
  if (warp == 0) tc::tmem_alloc<512>(&s_tmem);
  if (tid == 0) {
    tc::mbar_init(&s_bar, 1);
    tc::mbar_fence_init();
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const float v = B[i];
    const __half h = __float2half_rn(v);
    const size_t off = (size_t)(k >> 3) * N * 8 + (size_t)n * 8 + (k & 7);
    Bhi[off] = h;
    Blo[off] = __float2half_rn(v - __half2float(h));
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = s_tmem;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t colD = 0, colAhi = 256, colAlo = 256 + 80;
  // A row of this thread -> TMEM (hi and lo), 8 columns (= 16 halves = one K step) at a time
  for (int k0 = 0; k0 < K; k0 += 16) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tc::split2(A[(size_t)tid * K + k0 + 2 * j], A[(size_t)tid * K + k0 + 2 * j + 1], hi[j], lo[j]);
    tc::tmem_st8(tbase + lane_base + colAhi + k0 / 2, hi);
    tc::tmem_st8(tbase + lane_base + colAlo + k0 / 2, lo);
  }
  tc::tmem_wait_st();
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_f16(128, N, 0);
    const uint32_t bhi = tc::smem_u32(Bhi), blo = tc::smem_u32(Blo);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t dhi = tc::smem_desc_kmajor(bhi + ks * 2 * lbo, lbo, sbo);
      tc::mma_ts(tbase + colD, tbase + colAhi + ks * 8, dhi, idesc, ks > 0);
      if (split) {
        const uint64_t dlo = tc::smem_desc_kmajor(blo + ks * 2 * lbo, lbo, sbo);
        tc::mma_ts(tbase + colD, tbase + colAlo + ks * 8, dhi, idesc, true);
        tc::mma_ts(tbase + colD, tbase + colAhi + ks * 8, dlo, idesc, true);
      }
    }
    tc::mma_commit(&s_bar);
  }
  tc::mbar_wait(&s_bar, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tc::tmem_ld16(tbase + lane_base + colD + c0, r);
    tc::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tbase);
}

extern "C" int surf_tc_selftest(const float* d_A, const float* d_B, float* d_D, int32_t K, int32_t N, int32_t split,
                                void* stream) {
  SURF_CHECK_ARG(d_A && d_B && d_D, "null pointer");
  SURF_CHECK_ARG(K % 16 == 0 && K >= 16 && K <= 160, "K must be a multiple of 16 in [16,160]");
  SURF_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256, "N must be a multiple of 16 in [16,256]");
  const size_t smem = (size_t)2 * N * K * sizeof(__half);
  SURF_CUDA(cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tc_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(d_A, d_B, d_D, K, N, split);
  SURF_LAUNCH_CHECK();
  return 0;
}

