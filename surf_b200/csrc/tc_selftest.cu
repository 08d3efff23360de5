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

// ---------------------------------------------------------------------------------------------
// micro-benchmark: cycles for `reps` back-to-back tcgen05.mma (M=128, N, K=16), one CTA.
//   mode 0: TS, one accumulator (dependent chain)      mode 1: TS, two alternating accumulators
//   mode 2: SS, one accumulator                        mode 3: SS, two alternating accumulators
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_tc_bench(int N, int reps, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ __align__(8) uint64_t s_bar2[4];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc<512>(&s_tmem);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) tc::mbar_init(&s_bar2[i], i == 2 ? 1 : 1000000);
    tc::mbar_arrive(&s_bar2[2]);          // barrier 2 completes its phase 0 immediately
    tc::mbar_init(&s_bar, mode >= 10 ? 2 : (mode >= 4 ? mode - 3 : 1));
    tc::mbar_fence_init();
  }
  for (int i = tid; i < 16384; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = s_tmem;
  // mode >= 4: (mode - 3) issuer threads (lane 0 of warps 0..), each with its own accumulator (N <= 128), TS form
  const int n_issuers = mode >= 10 ? 2 : (mode >= 4 ? mode - 3 : 1);
  if ((tid & 31) == 0 && warp < n_issuers) {
    const uint32_t idesc = tc::idesc_f16(128, N, 0);
    const uint64_t d0 = tc::smem_desc_kmajor(tc::smem_u32(smem), (uint32_t)N * 16, 128);
    const uint32_t dlo = (uint32_t)d0, dhi = (uint32_t)(d0 >> 32);
    const uint64_t a0 = tc::smem_desc_kmajor(tc::smem_u32(smem) + 32768, 2048, 128);
    const uint32_t alo = (uint32_t)a0, ahi = (uint32_t)(a0 >> 32);
    const long long t0 = clock64();
    if (mode >= 10) {
      // mode 10: 2 issuers, commit to a scratch barrier after every 6 MMAs
      // mode 11: additionally wait on an (already completed) barrier before every group of 6
      // mode 12: mode 10 with distinct A/B addresses per MMA (like the real loop)
      for (int r = 0; r < reps; ++r) {
        const uint32_t tD = tbase + warp * 128;
        if (mode == 11 && (r % 6) == 0) tc::mbar_wait(&s_bar2[2], 0);
        const uint32_t off = (mode == 12) ? (uint32_t)(r % 6) * 64u : 0u;
        tc::mma_ts_w<true>(tD, tbase + 448 + (r & 1) * 8, dlo + off, dhi, idesc);
        if ((r % 6) == 5) tc::mma_commit(&s_bar2[warp]);
      }
    } else
    for (int r = 0; r < reps; ++r) {
      uint32_t tD = tbase + (((mode & 1) && mode < 4 && (r & 1)) ? 256 : 0);
      if (mode >= 4) tD = tbase + warp * 128;
      if (mode < 2 || mode >= 4) tc::mma_ts_w<true>(tD, tbase + 448, dlo, dhi, idesc);
      else tc::mma_ss_w<true>(tD, alo, ahi, dlo, dhi, idesc);
    }
    const long long t1 = clock64();
    tc::mma_commit(&s_bar);
    tc::mbar_wait(&s_bar, 0);
    const long long t2 = clock64();
    if (warp == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tbase);
}

extern "C" int surf_tc_bench(int32_t N, int32_t reps, int32_t mode, long long* h_out) {
  long long* d = nullptr;
  SURF_CUDA(cudaMalloc((void**)&d, 16));
  SURF_CUDA(cudaFuncSetAttribute(k_tc_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int it = 0; it < 2; ++it) k_tc_bench<<<1, 128, 65536>>>(N, reps, mode, d);
  SURF_CUDA(cudaDeviceSynchronize());
  SURF_CUDA(cudaMemcpy(h_out, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}
