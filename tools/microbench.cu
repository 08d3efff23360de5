// Micro-benchmarks of the tcgen05 issue rate and of epilogue variants (development tools, NOT part of the product
// library).  Build: tools/build_microbench.sh -> tools/libsurf_microbench.so, driven by tools/epi_bench.py / tc_bench.py.
#include "surf_internal.cuh"
#include "tc_common.cuh"

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
  SURF_LAUNCH_CHECK();
  SURF_CUDA(cudaDeviceSynchronize());
  SURF_CUDA(cudaMemcpy(h_out, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// micro-benchmark of the forward epilogue arithmetic (no TMEM): 16 warps, each thread runs `reps` groups of 8
// elements.  variant 0: softplus + e-code + sign + fp16 hi/lo split (the real thing); 1: MUFU only (ex2 + lg2);
// 2: everything but the MUFU ops; 3: softplus + split (no e-code / sign); 4: softplus only
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float eb_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float eb_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int VARIANT>
__global__ void __launch_bounds__(512, 1) k_epi_bench(int reps, float seed, long long* out, uint32_t* sink) {
  float z[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) z[n] = seed * (float)(threadIdx.x * 8 + n) - 0.3f;
  uint32_t acc = 0, sgn = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
    float h[8];
    uint32_t cw[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      float e, lg;
      if (VARIANT == 2) {
        e = fmaf(fabsf(z[n]), -0.25f, 0.9f);
        lg = fmaf(e, 0.7f, 0.1f);
      } else {
        e = eb_ex2(fabsf(z[n]) * -144.26950408889634f);
        lg = eb_lg2(1.0f + e);
      }
      h[n] = fmaf(lg, 0.0069314718055994531f, fmaxf(z[n], 0.f));
      if (VARIANT == 0 || VARIANT == 2) {
        cw[n] = __float_as_uint(fminf(e, 0.9999847412109375f) + 128.0f);
        sgn = __funnelshift_l(__float_as_uint(z[n]), sgn, 1);
      }
    }
    if (VARIANT == 0 || VARIANT == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc ^= __byte_perm(cw[2 * j], cw[2 * j + 1], 0x5410);
    }
    if (VARIANT == 0 || VARIANT == 2 || VARIANT == 3) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hi, lo;
        tc::split2(h[2 * j], h[2 * j + 1], hi, lo);
        acc ^= hi + lo;
      }
    } else {
#pragma unroll
      for (int n = 0; n < 8; ++n) acc ^= __float_as_uint(h[n]);
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) z[n] = z[n] * 0.999f + 1e-4f;      // next group's inputs
  }
  const long long t1 = clock64();
  sink[blockIdx.x * 512 + threadIdx.x] = acc ^ sgn;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}


// the same arithmetic with the TMEM / synchronisation traffic of the real epilogue added step by step:
//   LEVEL 1: + tcgen05.ld x8 of the group (wait::ld) and tcgen05.st x4 hi / lo   2: + tcgen05.wait::st
//   3: + tcgen05.fence::before_thread_sync + __syncwarp + elected mbarrier.arrive  4: + 16-byte code store to global
//   5: level 4 with the loads software-pipelined one group ahead
template <int LEVEL, int MMA>   // MMA: 0 none, 1 = a 17th warp issues TS-form MMAs (N = 128) all along, 2 = SS form,
                                // 3 = the 17th warp sits in mbar_wait (try_wait + suspend hint) all along, 4 = all 32 lanes of it do
__global__ void __launch_bounds__(544, 1) k_epi_bench2(int reps, long long* out, uint4* scratch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tc::tmem_alloc<512>(&s_tmem);
  if (threadIdx.x == 0) {
    tc::mbar_init(&s_bar, 1000000u);
    tc::mbar_fence_init();
  }
  __shared__ volatile int s_stop;
  __shared__ __align__(8) uint64_t s_bar2;
  if (threadIdx.x == 0) {
    s_stop = 0;
    tc::mbar_init(&s_bar2, 1);
    tc::mbar_fence_init();
  }
  if (MMA) for (int i = threadIdx.x; i < 16384; i += 544) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 16) {
    if (MMA >= 3) {
      if (MMA == 4 || lane == 0) tc::mbar_wait(&s_bar2, 0);
    } else if (MMA && lane == 0) {
      const uint32_t idesc = tc::idesc_f16(128, 128, 0);
      const uint64_t d0 = tc::smem_desc_kmajor(tc::smem_u32(smem), 2048, 128);
      const uint64_t a0 = tc::smem_desc_kmajor(tc::smem_u32(smem) + 32768, 2048, 128);
      long long n = 0;
      while (!s_stop) {
        for (int i = 0; i < 6; ++i) {
          if (MMA == 1) tc::mma_ts_w<true>(s_tmem + 160, s_tmem + 448, (uint32_t)d0, (uint32_t)(d0 >> 32), idesc);
          else tc::mma_ss_w<true>(s_tmem + 160, (uint32_t)a0, (uint32_t)(a0 >> 32), (uint32_t)d0, (uint32_t)(d0 >> 32), idesc);
        }
        tc::mma_commit(&s_bar);
        n += 6;
      }
      out[1] = n;
    }
    __syncthreads();
    return;
  }
  const uint32_t tl = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 8;
  uint32_t acc = 0, sgn = 0;
  uint32_t dn[8];
  {
    uint32_t init[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) init[n] = __float_as_uint(1e-3f * (float)(threadIdx.x * 8 + n) - 0.3f);
    for (int g = 0; g < 4; ++g) tc::tmem_st8(tl + g * 32, init);
    tc::tmem_wait_st();
  }
  asm volatile("bar.sync 1, 512;" ::: "memory");
  const long long t0 = clock64();
  if (LEVEL == 5) tc::tmem_ld8(tl, dn);
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
    const int g = r & 3;
    uint32_t dv[8];
    if (LEVEL == 5) {
      asm volatile("tcgen05.wait::ld.sync.aligned;"
                   : "+r"(dn[0]), "+r"(dn[1]), "+r"(dn[2]), "+r"(dn[3]), "+r"(dn[4]), "+r"(dn[5]), "+r"(dn[6]), "+r"(dn[7])::"memory");
#pragma unroll
      for (int n = 0; n < 8; ++n) dv[n] = dn[n];
      tc::tmem_ld8(tl + ((g + 1) & 3) * 32, dn);
    } else {
      tc::tmem_ld8(tl + g * 32, dv);
      tc::tmem_wait_ld();
    }
    float h[8];
    uint32_t cw[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float z = __uint_as_float(dv[n]);
      const float e = eb_ex2(fabsf(z) * -144.26950408889634f);
      h[n] = fmaf(eb_lg2(1.0f + e), 0.0069314718055994531f, fmaxf(z, 0.f));
      cw[n] = __float_as_uint(fminf(e, 0.9999847412109375f) + 128.0f);
      sgn = __funnelshift_l(__float_as_uint(z), sgn, 1);
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc::split2(h[2 * j], h[2 * j + 1], hi[j], lo[j]);
    tc::tmem_st4(tl + 320 + g * 16, hi);
    tc::tmem_st4(tl + 384 + g * 16, lo);
    if (LEVEL >= 2) tc::tmem_wait_st();
    if (LEVEL >= 3) {
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_bar);
    }
    const uint4 spw = make_uint4(__byte_perm(cw[0], cw[1], 0x5410), __byte_perm(cw[2], cw[3], 0x5410),
                                 __byte_perm(cw[4], cw[5], 0x5410), __byte_perm(cw[6], cw[7], 0x5410));
    if (LEVEL >= 4) scratch[(size_t)(blockIdx.x * 4 + g) * 512 + threadIdx.x] = spw;
    else acc ^= spw.x ^ spw.y ^ spw.z ^ spw.w;
  }
  if (LEVEL == 5) tc::tmem_wait_ld();
  const long long t1 = clock64();
  if (acc == 0x12345678u && sgn == 77u) scratch[threadIdx.x].x = acc + dn[0];
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  asm volatile("bar.sync 1, 512;" ::: "memory");
  if (threadIdx.x == 0) {
    s_stop = 1;
    tc::mbar_arrive(&s_bar2);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (MMA) {                      // drain the tensor pipe before giving TMEM back
      for (volatile int spin = 0; spin < 20000; ++spin) {}
    }
    tc::tmem_dealloc<512>(s_tmem);
  }
}

extern "C" int surf_epi_bench(int32_t variant, int32_t reps, long long* h_out) {
  long long* d = nullptr;
  uint32_t* sink = nullptr;
  SURF_CUDA(cudaMalloc((void**)&d, 16));
  SURF_CUDA(cudaMalloc((void**)&sink, 148 * 512 * 4 * 16));
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_epi_bench2<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_epi_bench2<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_epi_bench2<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_epi_bench2<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_epi_bench2<5, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_epi_bench2<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_epi_bench2<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_epi_bench2<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_epi_bench2<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    attr = true;
  }
  for (int it = 0; it < 2; ++it) {
    switch (variant) {
      case 0: k_epi_bench<0><<<148, 512>>>(reps, 1e-3f, d, sink); break;
      case 1: k_epi_bench<1><<<148, 512>>>(reps, 1e-3f, d, sink); break;
      case 2: k_epi_bench<2><<<148, 512>>>(reps, 1e-3f, d, sink); break;
      case 3: k_epi_bench<3><<<148, 512>>>(reps, 1e-3f, d, sink); break;
      case 4: k_epi_bench<4><<<148, 512>>>(reps, 1e-3f, d, sink); break;
      case 11: k_epi_bench2<1, 0><<<148, 544, 65536>>>(reps, d, (uint4*)sink); break;
      case 12: k_epi_bench2<2, 0><<<148, 544, 65536>>>(reps, d, (uint4*)sink); break;
      case 13: k_epi_bench2<3, 0><<<148, 544, 65536>>>(reps, d, (uint4*)sink); break;
      case 14: k_epi_bench2<4, 0><<<148, 544, 65536>>>(reps, d, (uint4*)sink); break;
      case 21: k_epi_bench2<4, 1><<<148, 544, 65536>>>(reps, d, (uint4*)sink); break;
      case 22: k_epi_bench2<4, 2><<<148, 544, 65536>>>(reps, d, (uint4*)sink); break;
      case 23: k_epi_bench2<4, 3><<<148, 544, 65536>>>(reps, d, (uint4*)sink); break;
      case 24: k_epi_bench2<4, 4><<<148, 544, 65536>>>(reps, d, (uint4*)sink); break;
      default: k_epi_bench2<5, 0><<<148, 544, 65536>>>(reps, d, (uint4*)sink); break;
    }
  }
  SURF_LAUNCH_CHECK();
  SURF_CUDA(cudaDeviceSynchronize());
  SURF_CUDA(cudaMemcpy(h_out, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  cudaFree(sink);
  return 0;
}
