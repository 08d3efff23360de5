// K2a + K3, tensor-core edition, one 128-point tile per CTA pass: forward + analytic input gradient.
//   reference: SDFNetworkSparse.sdf / .gradient (sdf_network.py:95-141), lookup_sparse_volume (projector.py:217-390)
//
// TMEM map (512 columns): two fp32 accumulators D_a [0,160) and D_b [160,320) — one per MMA-issuing thread, so the
// two threads needed to keep the tensor pipe fed (tools/tc_bench.py: one thread sustains one MMA per ~100-120 clk,
// the pipe takes ~69 clk for M128 N128 K16) never accumulate into the same tile and the result stays bitwise
// deterministic (the epilogue adds D_a + D_b) — and the fp16 hi / lo halves of the A operand [320,384) / [384,448).
// 16 epilogue warps: warp (q, part) owns TMEM lane quarter q (32 points) and 32 of the 128 columns.
// Forward: as sdf_tc.cu.  Reverse pass: delta_l (128 x 128) is the A operand, the transposed weights stream as
// N = 160 (128 hidden + 28 feature-gradient + 4 pad) x K = 32 chunks, softplus' is parked as unorm16 in a per-CTA
// L2-resident scratch, the feature-gradient columns are accumulated in registers across layers.
#include <math.h>
#include <string.h>

#include <vector>

#include "surf_internal.cuh"
#include "tc_common.cuh"

#define T1_EPI_WARPS 16
#define T1_EPI_THREADS (T1_EPI_WARPS * 32)
#define T1_THREADS ((T1_EPI_WARPS + 3) * 32)     // + 2 MMA issuers + 1 weight loader
#define T1_SLOT_BYTES 20480                      // N = 160 x K = 32 x (hi + lo)
#define T1_NSLOT 7

#define T1_DA 0u
#define T1_DB 160u
#define T1_AHI 320u
#define T1_ALO 384u

// dynamic smem (bytes)
#define S1_RING 0
#define S1_AFEAT (S1_RING + T1_NSLOT * T1_SLOT_BYTES)      // hi 8 KB | lo 8 KB  (128 rows x K 32)
#define S1_APE (S1_AFEAT + 16384)
#define S1_W6 (S1_APE + 16384)                             // 160 floats
#define S1_PART (S1_W6 + 640)                              // [4][128] floats
#define S1_GPE (S1_PART + 2048)                            // [28][128] floats
#define S1_GF (S1_GPE + 14336)                             // [28][128] floats ; later [4][3][128] partial grads
#define S1_BAR (S1_GF + 14336)
#define S1_TOTAL (S1_BAR + 256)

#ifdef TC_TRACE
// cheap in-kernel timeline: events go to a shared-memory log (a global atomic per event costs ~800 clk and distorts
// the single-threaded issuer), dumped to global memory by CTA 0 at kernel end.  Two writers: epilogue warp 0 lane 0
// (slot 0) and issuer 0 lane 0 (slot 1), 1024 events each.
__device__ long long g_t1_trace[8192];
__device__ int g_t1_trace_n;
#define T1_TRACE_SMEM 24576
#define TRACE1(ev)                                                                  \
  do {                                                                              \
    if (blockIdx.x == 0 && lane == 0 && trace_n < 768) {                           \
      long long* _t = reinterpret_cast<long long*>(smem + S1_TOTAL) + trace_slot * 1536; \
      _t[2 * trace_n] = (ev);                                                       \
      _t[2 * trace_n + 1] = clock64();                                              \
      trace_n++;                                                                    \
    }                                                                               \
  } while (0)
#else
#define T1_TRACE_SMEM 0
#define TRACE1(ev) do {} while (0)
#endif

struct T1Bars {
  uint64_t w_full[T1_NSLOT];
  uint64_t w_empty[T1_NSLOT];
  uint64_t d_full;
  uint64_t a_ready;
  uint32_t tmem_base;
};

__device__ __forceinline__ float t1_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float t1_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float t1_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// softplus(beta=100), branch-free (see sdf_tc.cu).  For the reverse pass the forward epilogue parks
// u = 1 + exp(-|100 z|) (16 mantissa bits, no int conversion: F2I / I2F / RCP all run on the quarter-rate XU
// pipe that already carries ex2 + lg2) plus the sign of z; softplus'(z) = sigmoid(100 z) = z >= 0 ? 1/u : 1 - 1/u
// is formed in the reverse epilogue.
__device__ __forceinline__ float t1_softplus(float z, float& u) {
  const float e = t1_ex2(fabsf(z) * -144.26950408889634f);
  u = 1.0f + e;
  return fmaf(t1_lg2(u), 0.0069314718055994531f, fmaxf(z, 0.f));
}
// u in [1,2] -> 16 bit code (round to nearest of the top 16 mantissa bits; 2.0 saturates) and back
__device__ __forceinline__ uint32_t t1_pack_u(float u) {
  // + 2^-17 (half a code step); genuine values saturate at code 0xfffe, 0xffff (+ sign) is reserved for PE columns
  const float c = fminf(u + 7.62939453125e-06f, 1.99997711181640625f);
  return (__float_as_uint(c) >> 7) & 0xffffu;
}
__device__ __forceinline__ float t1_unpack_u(uint32_t code) { return __uint_as_float(0x3f800000u | (code << 7)); }

// chunk layout per layer phase of the weight stream
__device__ __forceinline__ int t1_phase_chunks(int phase) {   // phases 0..5 fwd, 6..10 reverse lin5..lin1, 11 reverse lin0
  if (phase == 0) return 1;
  if (phase < 6) return 6;      // 4 x K32 hidden + 2 x K16 (features | bias)
  if (phase < 11) return 4;
  return 2;
}

struct EpiCtx {
  uint32_t tl;            // TMEM base of my lane quarter
  int c0, r, te;
  const uint8_t* ape;
  const float* sw6;
  float inv_scale;
  uint4* scratch;
  uint32_t* sgn_scratch;
  float* s_gpe;
};

__device__ __forceinline__ float t1_get_k(const uint8_t* base, int r, int k) {
  const uint32_t off = (uint32_t)(k >> 3) * 2048u + r * 16 + (k & 7) * 2;
  return __half2float(*reinterpret_cast<const __half*>(base + off)) +
         __half2float(*reinterpret_cast<const __half*>(base + 8192 + off));
}

// Forward epilogue of one hidden layer for my 32 columns.  HAS_B: a second accumulator exists (every layer but lin0);
// SKIP: my columns 101..127 become the positional encoding (input of the skip layer); HEAD: lin5 -> SDF head partial
// sum and (GRAD) delta5 = w6 / scale * softplus' written back as the first reverse A operand.
template <bool GRAD, bool HAS_B, bool SKIP, bool HEAD>
__device__ __forceinline__ void t1_fwd_epilogue(const EpiCtx& c, int l, float& head) {
  uint32_t sp[16];
  uint32_t sgn = 0;
#pragma unroll
  for (int hb = 0; hb < 2; ++hb) {
    const int cb = c.c0 + hb * 16;
    uint32_t a[16], b[16];
    tc::tmem_ld16(c.tl + T1_DA + cb, a);
    if (HAS_B) tc::tmem_ld16(c.tl + T1_DB + cb, b);
    tc::tmem_wait_ld();
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float z0 = __uint_as_float(a[2 * j]), z1 = __uint_as_float(a[2 * j + 1]);
      if (HAS_B) {
        z0 += __uint_as_float(b[2 * j]);
        z1 += __uint_as_float(b[2 * j + 1]);
      }
      float u0, u1;
      float h0 = t1_softplus(z0, u0);
      float h1 = t1_softplus(z1, u1);
      const int n0 = hb * 16 + 2 * j;                       // column offset inside my 32 (compile time)
      const bool pe0 = SKIP && (96 + n0 >= 101), pe1 = SKIP && (96 + n0 + 1 >= 101);   // SKIP => c0 == 96
      if (pe0) h0 = t1_get_k(c.ape, c.r, 96 + n0 - 101);
      if (pe1) h1 = t1_get_k(c.ape, c.r, 96 + n0 + 1 - 101);
      if (HEAD) {
        const float w0 = c.sw6[c.c0 + n0], w1 = c.sw6[c.c0 + n0 + 1];
        head = fmaf(h0, w0, head);
        head = fmaf(h1, w1, head);
        if (GRAD) {
          const float r0 = t1_rcp(u0), r1 = t1_rcp(u1);
          h0 = w0 * c.inv_scale * (z0 >= 0.f ? r0 : 1.0f - r0);
          h1 = w1 * c.inv_scale * (z1 >= 0.f ? r1 : 1.0f - r1);
        }
      } else if (GRAD) {
        // 16-bit code of u = 1 + exp(-|100 z|) and the sign of z; PE columns: reserved code 0xffff + sign set
        const uint32_t q0 = pe0 ? 0xffffu : t1_pack_u(u0), q1 = pe1 ? 0xffffu : t1_pack_u(u1);
        sp[hb * 8 + j] = q0 | (q1 << 16);
        if (pe0) sgn |= 1u << n0; else sgn |= (__float_as_uint(z0) >> 31) << n0;
        if (pe1) sgn |= 1u << (n0 + 1); else sgn |= (__float_as_uint(z1) >> 31) << (n0 + 1);
      }
      tc::split2(h0, h1, hi[j], lo[j]);
    }
    if (!HEAD || GRAD) {
      tc::tmem_st8(c.tl + T1_AHI + (cb >> 1), hi);
      tc::tmem_st8(c.tl + T1_ALO + (cb >> 1), lo);
    }
  }
  if (GRAD && !HEAD) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      c.scratch[(size_t)(l * 4 + j) * T1_EPI_THREADS + c.te] = make_uint4(sp[4 * j], sp[4 * j + 1], sp[4 * j + 2], sp[4 * j + 3]);
    c.sgn_scratch[(size_t)l * T1_EPI_THREADS + c.te] = sgn;
  }
}

// Reverse epilogue: delta_{l-1} = (D_a + D_b) * softplus'(z_{l-1}) for my 32 columns -> next A operand.
// SKIP_PE: this is the skip layer and my columns 101..127 are the PE input gradient (kept in smem, delta = 0).
template <bool SKIP_PE>
__device__ __forceinline__ void t1_bwd_epilogue(const EpiCtx& c, const uint4 (&spv)[4], uint32_t sgn) {
  const uint32_t* spw = reinterpret_cast<const uint32_t*>(spv);
#pragma unroll
  for (int hb = 0; hb < 2; ++hb) {
    const int cb = c.c0 + hb * 16;
    uint32_t a[16], b[16];
    tc::tmem_ld16(c.tl + T1_DA + cb, a);
    tc::tmem_ld16(c.tl + T1_DB + cb, b);
    tc::tmem_wait_ld();
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g0 = __uint_as_float(a[2 * j]) + __uint_as_float(b[2 * j]);
      const float g1 = __uint_as_float(a[2 * j + 1]) + __uint_as_float(b[2 * j + 1]);
      const int n0 = hb * 16 + 2 * j;
      const uint32_t w = spw[hb * 8 + j];
      const float r0 = t1_rcp(t1_unpack_u(w & 0xffffu)), r1 = t1_rcp(t1_unpack_u(w >> 16));
      float d0 = ((sgn >> n0) & 1u) ? 1.0f - r0 : r0;
      float d1 = ((sgn >> (n0 + 1)) & 1u) ? 1.0f - r1 : r1;
      if (SKIP_PE) {                                        // c0 == 96
        if (96 + n0 >= 101) { c.s_gpe[(96 + n0 - 101) * 128 + c.r] = g0; d0 = 0.f; }
        if (96 + n0 + 1 >= 101) { c.s_gpe[(96 + n0 + 1 - 101) * 128 + c.r] = g1; d1 = 0.f; }
      }
      tc::split2(g0 * d0, g1 * d1, hi[j], lo[j]);
    }
    tc::tmem_st8(c.tl + T1_AHI + (cb >> 1), hi);
    tc::tmem_st8(c.tl + T1_ALO + (cb >> 1), lo);
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(T1_THREADS, 1)
k_sdf_tc1(const DevScene sc, const DevNet net, const PointSource src, const uint8_t* __restrict__ wblob,
          const T1Stream stream, float* __restrict__ sdf_out, float* __restrict__ grad_out,
          uint4* __restrict__ scratch_all, int negate) {
  extern __shared__ __align__(1024) uint8_t smem[];
  T1Bars* bars = reinterpret_cast<T1Bars*>(smem + S1_BAR);
  float* sw6 = reinterpret_cast<float*>(smem + S1_W6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int trace_n = 0;
  const int trace_slot = (warp == 0) ? 0 : 1;
  (void)trace_n; (void)trace_slot;
  constexpr int NPHASE = GRAD ? 12 : 6;
  const int nch_tile = GRAD ? stream.n_all : stream.n_fwd;

  int64_t n_total = src.n;
  if (src.count) {
    const int64_t c = *src.count;
    n_total = c < n_total ? c : n_total;
  }
  const int64_t n_tiles = (n_total + 127) / 128;
  int64_t my_tiles = 0;
  if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == T1_EPI_WARPS) tc::tmem_alloc<512>(&bars->tmem_base);
  if (tid == 0) {
    for (int i = 0; i < T1_NSLOT; ++i) {
      tc::mbar_init(&bars->w_full[i], 1);
      tc::mbar_init(&bars->w_empty[i], 1);
    }
    tc::mbar_init(&bars->d_full, 2);
    tc::mbar_init(&bars->a_ready, T1_EPI_THREADS);
    tc::mbar_fence_init();
  }
  for (int i = tid; i < 160; i += T1_THREADS) sw6[i] = net.w6[i];
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = bars->tmem_base;

  if (warp < T1_EPI_WARPS) {
    // =============================== epilogue / staging warps ===============================
    const int q = warp & 3, part = warp >> 2;
    const int r = q * 32 + lane;                 // row = point = TMEM lane
    const int c0 = part * 32;                    // first of my 32 hidden columns
    const uint32_t tl = tbase + ((uint32_t)(q * 32) << 16);
    uint8_t* afeat = smem + S1_AFEAT;
    uint8_t* ape = smem + S1_APE;
    float* s_part = reinterpret_cast<float*>(smem + S1_PART);
    float* s_gpe = reinterpret_cast<float*>(smem + S1_GPE);
    float* s_gf = reinterpret_cast<float*>(smem + S1_GF);
    uint4* scratch = scratch_all + (size_t)blockIdx.x * (5 * 4 * T1_EPI_THREADS + 5 * T1_EPI_THREADS / 4);
    uint32_t* sgn_scratch = reinterpret_cast<uint32_t*>(scratch + 5 * 4 * T1_EPI_THREADS);
    const int te = warp * 32 + lane;             // 0..511
    uint32_t ph_d = 0;
    auto put_k = [&](uint8_t* base, int k, float v) {
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const uint32_t off = (uint32_t)(k >> 3) * 2048u + r * 16 + (k & 7) * 2;
      *reinterpret_cast<__half*>(base + off) = h;
      *reinterpret_cast<__half*>(base + 8192 + off) = l;
    };
    auto get_k = [&](const uint8_t* base, int k) {
      const uint32_t off = (uint32_t)(k >> 3) * 2048u + r * 16 + (k & 7) * 2;
      return __half2float(*reinterpret_cast<const __half*>(base + off)) +
             __half2float(*reinterpret_cast<const __half*>(base + 8192 + off));
    };
    auto epi_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(T1_EPI_THREADS) : "memory"); };
    auto load_point = [&](int64_t i, float& px, float& py, float& pz) -> int64_t {
      px = 0.f; py = 0.f; pz = 0.f;
      if (i >= n_total) return -1;
      const int64_t id = src.list ? (int64_t)src.list[i] : i;
      if (src.mode == 0) {
        px = src.pts[id * 3]; py = src.pts[id * 3 + 1]; pz = src.pts[id * 3 + 2];
      } else if (src.mode == 1) {
        const int64_t ray = id / src.S;
        const float t = src.mid_z[id];
        px = ray_at(src.rays_o[ray * 3], src.rays_d[ray * 3], t);
        py = ray_at(src.rays_o[ray * 3 + 1], src.rays_d[ray * 3 + 1], t);
        pz = ray_at(src.rays_o[ray * 3 + 2], src.rays_d[ray * 3 + 2], t);
      } else {
        const int64_t yz = (int64_t)src.ny * src.nz;
        const int xi = (int)(id / yz);
        const int rem = (int)(id - (int64_t)xi * yz);
        px = src.xs[xi]; py = src.ys[rem / src.nz]; pz = src.zs[rem % src.nz];
      }
      return id;
    };

    float nf7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // next tile's features of (row, level = part), gathered early
    bool have_next = false;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
      float px, py, pz;
      const int64_t id = load_point(tile * 128 + r, px, py, pz);
      // ---- staging: thread (row, part) gathers level `part` and encodes PE frequency `part` ----
      {
        float f7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (GRAD && have_next) {
#pragma unroll
          for (int c = 0; c < 7; ++c) f7[c] = nf7[c];
        } else if (part < sc.n_levels) {
          sparse_level<0>(sc, part, px, py, pz, nullptr, f7);
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) put_k(afeat, part * 7 + c, f7[c]);
        const float xs[3] = {px * net.scale, py * net.scale, pz * net.scale};
        if (part == 0) {
#pragma unroll
          for (int d = 0; d < 3; ++d) put_k(ape, d, xs[d]);
        }
        if (part == 3) {
          put_k(afeat, 28, 1.0f);
          put_k(ape, 27, 1.0f);
#pragma unroll
          for (int k = 29; k < 32; ++k) put_k(afeat, k, 0.f);
#pragma unroll
          for (int k = 28; k < 32; ++k) put_k(ape, k, 0.f);
        }
        const float fr = (float)(1 << part);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          float sn = 0.f, cs = 0.f;
          if (part < net.multires) sincosf(xs[d] * fr, &sn, &cs);
          put_k(ape, 3 + 6 * part + d, sn);
          put_k(ape, 3 + 6 * part + 3 + d, cs);
        }
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(&bars->a_ready);
      if (warp == 0) TRACE1(1);

      float gf[8];          // reverse pass: d sdf / d feat for feature columns part*8 .. part*8+7
      // ------------------------------------ forward ------------------------------------
      EpiCtx ec;
      ec.tl = tl; ec.c0 = c0; ec.r = r; ec.te = te; ec.ape = ape; ec.sw6 = sw6; ec.inv_scale = net.inv_scale;
      ec.scratch = scratch; ec.sgn_scratch = sgn_scratch; ec.s_gpe = s_gpe;
      for (int l = 0; l < 6; ++l) {
        if (l == 1 && it + 1 < my_tiles && part < sc.n_levels) {
          // while the tensor core works on this layer: pull the NEXT tile's gather working set into L2
          float nx, ny, nz;
          if (load_point((tile + gridDim.x) * 128 + r, nx, ny, nz) >= 0) sparse_prefetch_l2(sc, part, nx, ny, nz);
        }
        tc::mbar_wait(&bars->d_full, ph_d & 1);
        ph_d++;
        tc::tc_fence_after();
        if (warp == 0) TRACE1(10 + l);
        float head = 0.f;
        // layer-specialised epilogues: the generic hidden layer carries no special-case instructions
        if (l == 0) t1_fwd_epilogue<GRAD, false, false, false>(ec, l, head);
        else if (l == 5) t1_fwd_epilogue<GRAD, true, false, true>(ec, l, head);
        else if (l + 1 == net.skip_layer && part == 3) t1_fwd_epilogue<GRAD, true, true, false>(ec, l, head);
        else t1_fwd_epilogue<GRAD, true, false, false>(ec, l, head);
        if (l < 5 || GRAD) {
          tc::tmem_wait_st();
          tc::tc_fence_before();
          tc::mbar_arrive(&bars->a_ready);
          if (warp == 0) TRACE1(30 + l);
        }
        if (l == 5) {
          s_part[part * 128 + r] = head;
          epi_bar();
          if (part == 0) {
            float s = s_part[r] + s_part[128 + r] + s_part[256 + r] + s_part[384 + r] + net.b6;
#pragma unroll
            for (int c = 0; c < 28; ++c) s = fmaf(get_k(afeat, c), sw6[128 + c], s);
            s *= net.inv_scale;
            if (id >= 0) sdf_out[id] = (negate & 1) ? -s : s;
          }
          if (!GRAD) {
            tc::tc_fence_before();
            epi_bar();            // afeat / s_part reads done before the next tile restages
          }
        }
      }
      if (!GRAD) continue;

      // ------------------------------------ reverse ------------------------------------
#pragma unroll
      for (int j = 0; j < 8; ++j) gf[j] = sw6[128 + part * 8 + j] * net.inv_scale;
      for (int l = 5; l >= 1; --l) {
        // u codes / signs of layer l-1 for my columns (thread-private, written in the forward pass)
        uint4 spv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) spv[j] = scratch[(size_t)((l - 1) * 4 + j) * T1_EPI_THREADS + te];
        const uint32_t sgn = sgn_scratch[(size_t)(l - 1) * T1_EPI_THREADS + te];
        if (l == 4) {
          // gather the NEXT tile's features now (its lines were prefetched into L2 during lin1): the load latency
          // hides behind this layer's MMAs and the next tile's staging shrinks to the smem writes
          have_next = false;
          if (it + 1 < my_tiles) {
            float nx, ny, nz;
            load_point((tile + gridDim.x) * 128 + r, nx, ny, nz);
#pragma unroll
            for (int c = 0; c < 7; ++c) nf7[c] = 0.f;
            if (part < sc.n_levels) sparse_level<0>(sc, part, nx, ny, nz, nullptr, nf7);
            have_next = true;
          }
        }
        tc::mbar_wait(&bars->d_full, ph_d & 1);
        ph_d++;
        tc::tc_fence_after();
        if (warp == 0) TRACE1(16 + (5 - l));
        if (l == net.skip_layer && part == 3) t1_bwd_epilogue<true>(ec, spv, sgn);
        else t1_bwd_epilogue<false>(ec, spv, sgn);
        {   // feature-gradient columns 128 + part*8 ..
          uint32_t a[8], b[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
                       : "r"(tl + T1_DA + 128 + part * 8));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7])
                       : "r"(tl + T1_DB + 128 + part * 8));
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j) gf[j] += __uint_as_float(a[j]) + __uint_as_float(b[j]);
        }
        tc::tmem_wait_st();
        tc::tc_fence_before();
        tc::mbar_arrive(&bars->a_ready);
        if (warp == 0) TRACE1(36 + (5 - l));
      }
      // feature gradients -> smem (28 x 128)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (part * 8 + j < 28) s_gf[(part * 8 + j) * 128 + r] = gf[j];
      epi_bar();                                   // s_gpe (skip part) and s_gf complete
      // d feats / d x for level `part` (re-gather; lines were touched a tile ago, L2 hits) — issued before
      // waiting for the lin0 reverse MMAs so its latency overlaps them
      float o3[3] = {0.f, 0.f, 0.f};
      if (part < sc.n_levels) {
        float g7[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) g7[c] = s_gf[(part * 7 + c) * 128 + r];
        sparse_level<1>(sc, part, px, py, pz, g7, o3);
      }
      // ---- reverse of lin0: g_pe += delta0 . W0 (N = 32) ----
      tc::mbar_wait(&bars->d_full, ph_d & 1);
      ph_d++;
      tc::tc_fence_after();
      if (part == 0) {
        uint32_t a[16], b[16];
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          tc::tmem_ld16(tl + T1_DA + hb * 16, a);
          tc::tmem_ld16(tl + T1_DB + hb * 16, b);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = hb * 16 + j;
            if (k < 27) s_gpe[k * 128 + r] += __uint_as_float(a[j]) + __uint_as_float(b[j]);
          }
        }
      }
      tc::tc_fence_before();
      epi_bar();                                   // all s_gf reads done, s_gpe final
      float* s_pg = s_gf;                          // reuse as [4][3][128]
      s_pg[(part * 3 + 0) * 128 + r] = o3[0];
      s_pg[(part * 3 + 1) * 128 + r] = o3[1];
      s_pg[(part * 3 + 2) * 128 + r] = o3[2];
      epi_bar();
      if (part < 3 && id >= 0) {                   // thread (row, d = part) finishes component d
        const int d = part;
        float gx = s_gpe[d * 128 + r];
        float fr = 1.0f;
        for (int f = 0; f < net.multires; ++f) {
          const float sn = get_k(ape, 3 + 6 * f + d), cs = get_k(ape, 3 + 6 * f + 3 + d);
          gx += fr * (s_gpe[(3 + 6 * f + d) * 128 + r] * cs - s_gpe[(3 + 6 * f + 3 + d) * 128 + r] * sn);
          fr *= 2.0f;
        }
        gx *= net.scale;
#pragma unroll
        for (int lv = 0; lv < 4; ++lv) gx += s_pg[(lv * 3 + d) * 128 + r];
        grad_out[id * 3 + d] = gx;
      }
      epi_bar();                                   // smem scratch free for the next tile
      if (warp == 0) TRACE1(99);
    }
  } else if (warp < T1_EPI_WARPS + 2) {
    // =============================== MMA issuers: sub 0 -> D_a, sub 1 -> D_b ===============================
    const int sub = warp - T1_EPI_WARPS;
    const bool fast = (negate & 2) != 0;      // single fp16 MMA per product (opt-in reduced-precision mode)
    if (tc::elect_one()) {
      const uint32_t ring = tc::smem_u32(smem + S1_RING);
      const uint32_t tD = tbase + (sub ? T1_DB : T1_DA);
      const uint32_t tAhi = tbase + T1_AHI, tAlo = tbase + T1_ALO;
      const uint32_t id128 = tc::idesc_f16(128, 128, 0), id160 = tc::idesc_f16(128, 160, 0), id32 = tc::idesc_f16(128, 32, 0);
      // descriptor constant parts: SBO 128; LBO = rows * 16
      const uint64_t d128 = tc::smem_desc_kmajor(0, 2048, 128), d160 = tc::smem_desc_kmajor(0, 2560, 128),
                     d32 = tc::smem_desc_kmajor(0, 512, 128);
      const uint32_t afeat_lo = (uint32_t)d128 | (tc::smem_u32(smem + S1_AFEAT) >> 4);
      const uint32_t ape_lo = (uint32_t)d128 | (tc::smem_u32(smem + S1_APE) >> 4);
      uint32_t ph_a = 0;
      int slot = 0;            // ring position / parity kept incrementally (a 64-bit % per chunk costs ~400 clk)
      uint32_t ring_par = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        for (int p = 0; p < NPHASE; ++p) {
          const int nch = t1_phase_chunks(p);
          tc::mbar_wait(&bars->a_ready, ph_a & 1);
          ph_a++;
          tc::tc_fence_after();
          if (sub == 0) TRACE1(50 + p);
          for (int c = 0; c < nch; ++c, slot = (slot + 1 == T1_NSLOT) ? 0 : slot + 1, ring_par ^= (slot == 0)) {
            if ((c & 1) != sub) continue;
            if (sub == 0) TRACE1(1000 + p * 8 + c);
            tc::mbar_wait(&bars->w_full[slot], ring_par);
            if (sub == 0) TRACE1(2000 + p * 8 + c);
            const uint32_t wa = (ring + slot * T1_SLOT_BYTES) >> 4;       // 16-byte units
            const bool first = (c == sub);                                  // first chunk of this accumulator
            if (p < 6) {
              // forward chunk: N = 128, K = 32 ; hi at +0, lo at +8192 B ; K step = 2 groups = 4096 B
              const uint32_t w0 = (uint32_t)d128 | wa, dh = (uint32_t)(d128 >> 32);
              if (p == 0) {
                const uint32_t a0 = ape_lo;
                if (first) tc::mma_ss_w<false>(tD, a0, dh, w0, dh, id128); else tc::mma_ss_w<true>(tD, a0, dh, w0, dh, id128);
                if (!fast) tc::mma_ss_w<true>(tD, a0 + 512, dh, w0, dh, id128);
                if (!fast) tc::mma_ss_w<true>(tD, a0, dh, w0 + 512, dh, id128);
                tc::mma_ss_w<true>(tD, a0 + 256, dh, w0 + 256, dh, id128);
                if (!fast) tc::mma_ss_w<true>(tD, a0 + 768, dh, w0 + 256, dh, id128);
                if (!fast) tc::mma_ss_w<true>(tD, a0 + 256, dh, w0 + 768, dh, id128);
              } else if (c >= 4) {
                // half chunk (K = 16): feature columns; chunk 4 -> K step 0, chunk 5 -> K step 1 of the smem A operand
                const uint32_t a0 = afeat_lo + (c - 4) * 256;
                tc::mma_ss_w<true>(tD, a0, dh, w0, dh, id128);
                if (!fast) tc::mma_ss_w<true>(tD, a0 + 512, dh, w0, dh, id128);
                if (!fast) tc::mma_ss_w<true>(tD, a0, dh, w0 + 256, dh, id128);
              } else {
                const uint32_t ah = tAhi + c * 16, al = tAlo + c * 16;
                if (first) tc::mma_ts_w<false>(tD, ah, w0, dh, id128); else tc::mma_ts_w<true>(tD, ah, w0, dh, id128);
                if (!fast) tc::mma_ts_w<true>(tD, al, w0, dh, id128);
                if (!fast) tc::mma_ts_w<true>(tD, ah, w0 + 512, dh, id128);
                tc::mma_ts_w<true>(tD, ah + 8, w0 + 256, dh, id128);
                if (!fast) tc::mma_ts_w<true>(tD, al + 8, w0 + 256, dh, id128);
                if (!fast) tc::mma_ts_w<true>(tD, ah + 8, w0 + 768, dh, id128);
              }
            } else if (p < 11) {
              // reverse chunk: N = 160, K = 32 ; hi at +0, lo at +10240 B ; K step = 2 groups = 5120 B
              const uint32_t w0 = (uint32_t)d160 | wa, dh = (uint32_t)(d160 >> 32);
              const uint32_t ah = tAhi + c * 16, al = tAlo + c * 16;
              if (first) tc::mma_ts_w<false>(tD, ah, w0, dh, id160); else tc::mma_ts_w<true>(tD, ah, w0, dh, id160);
              if (!fast) tc::mma_ts_w<true>(tD, al, w0, dh, id160);
              if (!fast) tc::mma_ts_w<true>(tD, ah, w0 + 640, dh, id160);
              tc::mma_ts_w<true>(tD, ah + 8, w0 + 320, dh, id160);
              if (!fast) tc::mma_ts_w<true>(tD, al + 8, w0 + 320, dh, id160);
              if (!fast) tc::mma_ts_w<true>(tD, ah + 8, w0 + 960, dh, id160);
            } else {
              // reverse of lin0: N = 32, K = 64 per chunk (4 K steps) ; hi at +0, lo at +4096 B ; K step = 1024 B
              const uint32_t w0 = (uint32_t)d32 | wa, dh = (uint32_t)(d32 >> 32);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t ah = tAhi + c * 32 + ks * 8, al = tAlo + c * 32 + ks * 8;
                if (first && ks == 0) tc::mma_ts_w<false>(tD, ah, w0, dh, id32); else tc::mma_ts_w<true>(tD, ah, w0 + ks * 64, dh, id32);
                if (!fast) tc::mma_ts_w<true>(tD, al, w0 + ks * 64, dh, id32);
                if (!fast) tc::mma_ts_w<true>(tD, ah, w0 + 256 + ks * 64, dh, id32);
              }
            }
            tc::mma_commit(&bars->w_empty[slot]);
            if (sub == 0) TRACE1(3000 + p * 8 + c);
          }
          tc::mma_commit(&bars->d_full);
          if (sub == 0) TRACE1(70 + p);
        }
      }
    }
  } else {
    // =============================== weight loader ===============================
    if (lane == 0) {
      const int64_t total = my_tiles * nch_tile;
      int slot = 0, cid = 0;
      uint32_t par = 1;          // parity of the previous use of this slot (first round: nothing to wait for)
      for (int64_t s = 0; s < total; ++s, slot = (slot + 1 == T1_NSLOT) ? 0 : slot + 1, par ^= (slot == 0),
                   cid = (cid + 1 == nch_tile) ? 0 : cid + 1) {
        if (s >= T1_NSLOT) tc::mbar_wait(&bars->w_empty[slot], par);
        tc::mbar_arrive_expect_tx(&bars->w_full[slot], stream.bytes[cid]);
        tc::bulk_g2s(smem + S1_RING + slot * T1_SLOT_BYTES, wblob + stream.off[cid], stream.bytes[cid],
                     &bars->w_full[slot]);
      }
    }
  }
#ifdef TC_TRACE
  if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == T1_EPI_WARPS)) {
    const long long* _t = reinterpret_cast<const long long*>(smem + S1_TOTAL) + trace_slot * 1536;
    const int base = atomicAdd(&g_t1_trace_n, trace_n);
    for (int i = 0; i < trace_n && base + i < 4096; ++i) {
      g_t1_trace[2 * (base + i)] = _t[2 * i];
      g_t1_trace[2 * (base + i) + 1] = _t[2 * i + 1];
    }
  }
#endif
  tc::tc_fence_before();
  __syncthreads();
  if (warp == T1_EPI_WARPS) tc::tmem_dealloc<512>(tbase);
}

// ---------------------------------------------------------------------------------------------
// host: combined forward + reverse weight stream
// ---------------------------------------------------------------------------------------------
static inline uint16_t t1_f2h(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
static inline float t1_h2f(uint16_t b) {
  __half h;
  memcpy(&h, &b, 2);
  return __half2float(h);
}

T1Stream g_t1_stream;            // identical for every net of the supported shape; filled at build time

int surf_build_tc1_weights(const std::vector<std::vector<float>>& W, const surf_net_inputs* in, surf_net* net,
                           cudaStream_t st, int (*dev_alloc)(surf_net*, void**, size_t)) {
  T1Stream& S = g_t1_stream;
  memset(&S, 0, sizeof(S));
  std::vector<uint16_t> blob;
  int nc = 0;
  // element (n, kk) of a chunk with `rows` N-rows: hi half at [0, half_bytes), lo half after it
  auto add_chunk = [&](int rows, int K) {
    const size_t half = (size_t)rows * K;           // halves
    S.off[nc] = (uint32_t)(blob.size() * 2);
    S.bytes[nc] = (uint32_t)(half * 2 * 2);
    blob.resize(blob.size() + half * 2, 0);
    return blob.size() - half * 2;
  };
  auto put = [&](size_t base, int rows, int K, int n, int kk, float v) {
    const uint16_t hi = t1_f2h(v);
    const uint16_t lo = t1_f2h(v - t1_h2f(hi));
    const size_t off = (size_t)(kk >> 3) * rows * 8 + (size_t)n * 8 + (kk & 7);
    blob[base + off] = hi;
    blob[base + (size_t)rows * K + off] = lo;
  };
  // forward: lin0 (K = 27 + bias), lin1..lin5 (5 chunks of K = 32: 128 hidden | 28 feats + bias + pad)
  {
    const int O = in->out_dim[0], I = in->in_dim[0];
    const size_t b = add_chunk(128, 32);
    for (int n = 0; n < O && n < 128; ++n) {
      for (int k = 0; k < I; ++k) put(b, 128, 32, n, k, W[0][(size_t)n * I + k]);
      put(b, 128, 32, n, 27, in->h_bias[0][n]);
    }
    nc++;
  }
  for (int l = 1; l < 6; ++l) {
    const int O = in->out_dim[l], I = in->in_dim[l];
    for (int c = 0; c < 6; ++c) {
      const int K = c < 4 ? 32 : 16;                  // 4 hidden chunks, then the feature / bias columns in two halves
      const int kbase = c < 4 ? c * 32 : 128 + (c - 4) * 16;
      const size_t b = add_chunk(128, K);
      for (int n = 0; n < O && n < 128; ++n)
        for (int kk = 0; kk < K; ++kk) {
          const int k = kbase + kk;
          if (k < I) put(b, 128, K, n, kk, W[l][(size_t)n * I + k]);
          else if (k == 156) put(b, 128, K, n, kk, in->h_bias[l][n]);
        }
      nc++;
    }
  }
  S.n_fwd = nc;
  // reverse lin5..lin1: B[n = input index (160 rows)][k = output index], 4 chunks of K = 32
  for (int l = 5; l >= 1; --l) {
    const int O = in->out_dim[l], I = in->in_dim[l];
    for (int c = 0; c < 4; ++c) {
      const size_t b = add_chunk(160, 32);
      for (int kk = 0; kk < 32; ++kk) {
        const int k = c * 32 + kk;
        if (k >= O) continue;
        for (int n = 0; n < I && n < 160; ++n) put(b, 160, 32, n, kk, W[l][(size_t)k * I + n]);
      }
      nc++;
    }
  }
  // reverse lin0: B[n = PE index (32 rows)][k = output index], 2 chunks of K = 64
  {
    const int O = in->out_dim[0], I = in->in_dim[0];
    for (int c = 0; c < 2; ++c) {
      const size_t b = add_chunk(32, 64);
      for (int kk = 0; kk < 64; ++kk) {
        const int k = c * 64 + kk;
        if (k >= O) continue;
        for (int n = 0; n < I && n < 32; ++n) put(b, 32, 64, n, kk, W[0][(size_t)k * I + n]);
      }
      nc++;
    }
  }
  S.n_all = nc;
  void* p = nullptr;
  int rc = dev_alloc(net, &p, blob.size() * 2);
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice, st));
  SURF_CUDA(cudaStreamSynchronize(st));
  net->tc1_blob = (const uint8_t*)p;
  // softplus' scratch: 5 layers x 4 x 512 threads x 16 B per CTA
  // (x 2: sdf_tc3.cu keeps two tiles in flight per CTA)
  rc = dev_alloc(net, &p, (size_t)net->n_sm * 2 * (5 * 4 * T1_EPI_THREADS + 5 * T1_EPI_THREADS / 4) * sizeof(uint4));
  if (rc) return rc;
  net->tc1_scratch = p;
  return 0;
}

int launch_sdf_tc1(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_sdf, float* d_grad,
                   bool negate, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SURF_CUDA(cudaFuncSetAttribute(k_sdf_tc1<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S1_TOTAL + T1_TRACE_SMEM));
    SURF_CUDA(cudaFuncSetAttribute(k_sdf_tc1<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S1_TOTAL + T1_TRACE_SMEM));
    attr_set = true;
  }
  if (src.n <= 0) return 0;
  const int64_t tiles = (src.n + 127) / 128;
  const int grid = (int)(tiles < n->n_sm ? tiles : n->n_sm);
  surf_time_begin(d_grad ? 0 : 1, st);
  if (d_grad) {
    k_sdf_tc1<true><<<grid, T1_THREADS, S1_TOTAL + T1_TRACE_SMEM, st>>>(s->dev, n->dev, src, n->tc1_blob, g_t1_stream, d_sdf, d_grad,
                                                        (uint4*)n->tc1_scratch, (negate ? 1 : 0) | (surf_mlp_mode() == 4 ? 2 : 0));
  } else {
    k_sdf_tc1<false><<<grid, T1_THREADS, S1_TOTAL + T1_TRACE_SMEM, st>>>(s->dev, n->dev, src, n->tc1_blob, g_t1_stream, d_sdf, nullptr,
                                                         (uint4*)n->tc1_scratch, (negate ? 1 : 0) | (surf_mlp_mode() == 4 ? 2 : 0));
  }
  surf_time_end(d_grad ? 0 : 1, st);
  SURF_LAUNCH_CHECK();
  return 0;
}

#ifdef TC_TRACE
extern "C" int surf_t1_trace_read(long long* h_out, int max_events) {
  int n = 0;
  cudaMemcpyFromSymbol(&n, g_t1_trace_n, sizeof(int));
  if (n > max_events) n = max_events;
  if (n > 4096) n = 4096;
  cudaMemcpyFromSymbol(h_out, g_t1_trace, sizeof(long long) * 2 * n);
  int zero = 0;
  cudaMemcpyToSymbol(g_t1_trace_n, &zero, sizeof(int));
  return n;
}
#endif
