// k_sdf_tc4: sparse gather + SDF MLP forward + input gradient with TWO 128-point tiles per CTA half a step apart.
//
// k_sdf_tc2 (the kernel this one replaces on the render path) keeps one tile per SM: per layer the chain
//   TMEM load -> activation -> TMEM store -> fence -> mbarrier -> MMA issue -> commit
// is serial, and neither fewer epilogue instructions nor less MUFU work moved its time (tools/experiments/README.md).
// Here the 16 epilogue warps alternate between two tiles A and B: while they run the activation epilogue of A's layer
// the tensor core runs B's, and vice versa:
//     issuer  : MMA(A,s) | MMA(B,s) | load W(s+1) | MMA(A,s+1) | MMA(B,s+1) | ...
//     epilogue:          | EPI(A,s) | EPI(B,s)                 | EPI(A,s+1) | ...
// One weight buffer serves both tiles (832 KB from L2 per PAIR of tiles instead of per tile).  Layout as in
// k_sdf_smooth_tc (same weight blob, same TMEM budget): tile A's A operand in TMEM (TS form), tile B's in shared memory
// (SS form), accumulators D_A / D_B 128 columns each, the 28 feature-gradient columns of both tiles in persistent
// N = 32 accumulators F_A / F_B that the tensor core sums over the reverse layers.  fp16 hi/lo split, 3 MMAs per product.
// Thread (r, cq): point r of the current tile, columns cq*32 .. cq*32+31 of a layer, sparse level cq at both ends.
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include "smooth_common.cuh"
#include "surf_internal.cuh"
#include "tc_common.cuh"
#include "tc_layer_common.cuh"

#define T4_EPI_WARPS 16
#define T4_THREADS ((T4_EPI_WARPS + 1) * 32)
// TMEM columns
#define T4_AP_HI 0u
#define T4_AP_LO 64u
#define T4_D(x) (128u + 128u * (x))
#define T4_F(x) (384u + 32u * (x))
// shared memory
#define T4_SM_W 0
#define T4_SM_AB ST_W_BYTES                     // tile B's A operand: hi 32 KB | lo 32 KB
#define T4_SM_AF(x) (T4_SM_AB + 65536 + 16384 * (x))   // feature operand of tile x (K = 32): hi 8 KB | lo 8 KB
#define T4_SM_PART (T4_SM_AB + 65536 + 32768)   // float part[128][4][4]: partial (grad xyz, sdf head) per column quarter
#define T4_SM_PE (T4_SM_PART + 128 * 16 * 4)    // float PE[2][27][128]: positional encoding of both tiles
#define T4_SM_BAR (T4_SM_PE + 2 * 27 * 128 * 4)
#define T4_SM_TOTAL (T4_SM_BAR + 64)
// per CTA: softplus' codes of layers 0..4 of both tiles, one 32-bit word per column pair: e = exp(-|100 z|) as fp16 with
// the sign of z in the (otherwise unused) sign bit — the code of sdf_tc2.cu; 320 KB per CTA stays L2-resident where
// fp32 derivatives (640 KB per CTA, 97 MB in all) spilled to HBM: 9 long-scoreboard stalls per issue
#define T4_SCRATCH_WORDS (2 * 5 * 64 * 128)

static_assert(T4_SM_TOTAL <= 227 * 1024, "shared memory of k_sdf_tc4");
static_assert(T4_SCRATCH_WORDS * 4 <= 5 * 128 * 128 * 8, "k_sdf_tc4 shares the scratch of k_sdf_smooth_tc");

// softplus(beta = 100) and e = exp(-|100 z|)
__device__ __forceinline__ float t4_softplus(float z, float& e) {
  e = st_ex2(fabsf(z) * -144.26950408889634f);
  return fmaf(st_lg2(1.0f + e), 0.0069314718055994531f, fmaxf(z, 0.f));
}
__device__ __forceinline__ uint32_t t4_code_pair(float e0, float z0, float e1, float z1) {
  const __half2 h = __floats2half2_rn(e0, e1);
  return *reinterpret_cast<const uint32_t*>(&h) | ((__float_as_uint(z0) >> 16) & 0x8000u) | (__float_as_uint(z1) & 0x80000000u);
}
// softplus'(z) of the two columns of a code word
__device__ __forceinline__ void t4_decode_pair(uint32_t w, float& d0, float& d1) {
  const uint32_t m = w & 0x7fff7fffu;
  const float r0 = __fdividef(1.0f, tc::add_half<0>(m, 1.0f)), r1 = __fdividef(1.0f, tc::add_half<1>(m, 1.0f));
  d0 = (w & 0x8000u) ? 1.0f - r0 : r0;
  d1 = (w & 0x80000000u) ? 1.0f - r1 : r1;
}

struct T4Bars {
  uint64_t a_ready[2];   // the A operand of tile x for its next step is in place: one arrival per epilogue warp
  uint64_t d_full[2];    // the MMAs of tile x's step are done
  uint64_t w_full;       // the weights of a step have landed
  uint32_t tmem_base;
};

// positional encoding of one coordinate set: pe[27] (x3, then per frequency sin3 cos3)
__device__ __forceinline__ void t4_pe(float px, float py, float pz, float scale, float (&pe)[32]) {
#pragma unroll
  for (int k = 0; k < 32; ++k) pe[k] = 0.f;
  const float x[3] = {px, py, pz};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float X = x[d] * scale;
    pe[d] = X;
    float fr = 1.0f;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      float sn, cs;
      sincosf(X * fr, &sn, &cs);
      pe[3 + 6 * f + d] = sn;
      pe[3 + 6 * f + 3 + d] = cs;
      fr *= 2.0f;
    }
  }
}
// d sdf / d X_d (per unit of the scaled coordinate) from the gradient w.r.t. PE input j; PE: [27][128] of my tile
__device__ __forceinline__ void t4_pe_accum(const float* PE, int r, int j, float g, float (&g1)[3]) {
  if (j < 3) {
    g1[j] += g;
    return;
  }
  const int t = j - 3, f = t / 6, rem = t % 6, d = rem % 3;
  const float fr = (float)(1 << f);
  if (rem < 3) g1[d] = fmaf(fr * g, PE[(3 + 6 * f + 3 + d) * 128 + r], g1[d]);       // sin: fr cos
  else g1[d] = fmaf(-fr * g, PE[(3 + 6 * f + d) * 128 + r], g1[d]);                   // cos: -fr sin
}

struct T4Point {
  float px, py, pz;
  int32_t id;          // output index, -1: row beyond the list
};

__device__ __forceinline__ T4Point t4_point(const PointSource& src, int64_t i, int64_t n_total) {
  T4Point P;
  P.px = P.py = P.pz = 0.f;
  P.id = -1;
  if (i >= n_total) return P;
  if (src.mode == 0) {
    P.id = (int32_t)i;
    P.px = src.pts[i * 3]; P.py = src.pts[i * 3 + 1]; P.pz = src.pts[i * 3 + 2];
  } else {
    const int64_t id = src.list ? (int64_t)src.list[i] : i;
    const int64_t ray = id / src.S;
    const float t = src.mid_z[id];
    P.id = (int32_t)id;
    P.px = ray_at(src.rays_o[ray * 3], src.rays_d[ray * 3], t);
    P.py = ray_at(src.rays_o[ray * 3 + 1], src.rays_d[ray * 3 + 1], t);
    P.pz = ray_at(src.rays_o[ray * 3 + 2], src.rays_d[ray * 3 + 2], t);
  }
  return P;
}

__global__ void __launch_bounds__(T4_THREADS, 1)
k_sdf_tc4(const DevScene sc, const DevNet net, const PointSource src, const uint8_t* __restrict__ wblob,
          float* __restrict__ sdf_out, float* __restrict__ grad_out, uint32_t* __restrict__ scratch_all, int flags_i) {
  const bool fast = (flags_i & 2) != 0, negate = (flags_i & 1) != 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  T4Bars* bars = reinterpret_cast<T4Bars*>(smem + T4_SM_BAR);
  float* PART = reinterpret_cast<float*>(smem + T4_SM_PART);
  float* PEs = reinterpret_cast<float*>(smem + T4_SM_PE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // feature operands: zero; ones column (bias row of the weights) at k = 28
  for (int i = tid; i < 32768 / 16; i += T4_THREADS) reinterpret_cast<uint4*>(smem + T4_SM_AF(0))[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == T4_EPI_WARPS) tc::tmem_alloc<512>(&bars->tmem_base);
  if (tid == 0) {
    for (int x = 0; x < 2; ++x) {
      tc::mbar_init(&bars->a_ready[x], T4_EPI_WARPS);
      tc::mbar_init(&bars->d_full[x], 1);
    }
    tc::mbar_init(&bars->w_full, 1);
    tc::mbar_fence_init();
  }
  __syncthreads();
  if (tid < 256) *reinterpret_cast<__half*>(smem + T4_SM_AF(tid >> 7) + 3 * 2048 + (tid & 127) * 16 + 4 * 2) = __float2half_rn(1.0f);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = bars->tmem_base;

  int64_t n_total = src.n;
  if (src.count) {
    const int64_t c = *src.count;
    n_total = c < n_total ? c : n_total;
  }
  const int64_t n_tiles = (n_total + ST_ROWS - 1) / ST_ROWS;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const float scale = net.scale, c6s = net.inv_scale;

  if (warp < T4_EPI_WARPS) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t tl = tbase + ((uint32_t)(q * 32) << 16);
    uint32_t* scratch = scratch_all + (size_t)blockIdx.x * T4_SCRATCH_WORDS;
    const int quad_bar = 1 + q;
    auto quad_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(quad_bar) : "memory"); };
    uint32_t ph[2] = {0u, 0u};
    auto signal = [&](int x) {
      tc::tmem_wait_st();
      tc::fence_proxy_async();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->a_ready[x]);
    };
    auto wait_d = [&](int x) {
      tc::mbar_wait(&bars->d_full[x], ph[x] & 1);
      ph[x]++;
      tc::tc_fence_after();
    };
    // my 16 columns cb .. cb+15 of the next A operand of tile x
    auto store_a = [&](int x, int cb, const float (&a)[16]) {
      if (x == 0) st_store_tmem(tl + T4_AP_HI, tl + T4_AP_LO, cb, a);
      else st_store_smem(smem + T4_SM_AB, 32768, r, cb, a);
    };

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      T4Point P[2];
      float head[2];          // my part of the SDF head: features of my level + my 32 columns of lin5's output
      float g1[2][3];         // (column quarter 3 / 0) d sdf / d X from the PE columns of the skip layer / of lin0
      // ---- inputs of both tiles: positional encoding (column quarter 0), sparse features of level cq ----
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        P[x] = t4_point(src, (pair * 2 + x) * ST_ROWS + r, n_total);
        g1[x][0] = g1[x][1] = g1[x][2] = 0.f;
        if (cq == 0) {
          float pe[32];
          t4_pe(P[x].px, P[x].py, P[x].pz, scale, pe);
#pragma unroll
          for (int k = 0; k < 27; ++k) PEs[(x * 27 + k) * 128 + r] = pe[k];
          pe[27] = 1.0f;                 // bias row of lin0
          float a[16], b[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) { a[k] = pe[k]; b[k] = pe[16 + k]; }
          store_a(x, 0, a);
          store_a(x, 16, b);
        }
        float f7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (P[x].id >= 0 && cq < sc.n_levels) sparse_value_batched(sc, cq, P[x].px, P[x].py, P[x].pz, f7);
        float hp = 0.f;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          st_put_half(smem + T4_SM_AF(x), 8192, r, cq * 7 + c, f7[c]);
          hp = fmaf(f7[c], net.w6[128 + cq * 7 + c], hp);
        }
        head[x] = hp;
        signal(x);
      }
      quad_sync();            // the positional encodings of my points are visible to the other column quarters
      // ---- forward: lin0 .. lin5, the two tiles alternating ----
#pragma unroll 1
      for (int l = 0; l < 6; ++l) {
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          wait_d(x);
          uint32_t* sl = scratch + (size_t)((x * 5 + l) * 64 + cq * 16) * 128 + r;
          const bool pe_cols = (l == 2) && (cq == 3);     // the skip layer's input: columns 101..127 are the positional encoding
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int cb = cq * 32 + c * 16;
            uint32_t z[16];
            tc::tmem_ld16(tl + T4_D(x) + cb, z);
            tc::tmem_wait_ld();
            float a[16];
            if (l < 5) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float z0 = __uint_as_float(z[j]), z1 = __uint_as_float(z[j + 1]);
                float e0, e1;
                a[j] = t4_softplus(z0, e0);
                a[j + 1] = t4_softplus(z1, e1);
                sl[(size_t)(c * 8 + (j >> 1)) * 128] = t4_code_pair(e0, z0, e1, z1);
              }
            } else {         // SDF head; delta_5 = ga_6 s'(z_5)
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float h, d1, d2;
                st_softplus(__uint_as_float(z[j]), h, d1, d2);
                const float w = net.w6[cb + j];
                head[x] = fmaf(h, w, head[x]);
                a[j] = w * c6s * d1;
              }
            }
            if (pe_cols) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int k = 96 + c * 16 + j - 101;        // compile-time
                if (k >= 0) a[j] = PEs[(x * 27 + k) * 128 + r];
              }
            }
            store_a(x, cb, a);
          }
          signal(x);
        }
      }
      // ---- reverse: lin5 .. lin1 ----
#pragma unroll 1
      for (int l = 5; l >= 1; --l) {
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const uint32_t* sl = scratch + (size_t)((x * 5 + l - 1) * 64 + cq * 16) * 128 + r;
          uint32_t cw[16];        // softplus' codes of layer l-1 at my 32 columns, fetched before the accumulator wait
#pragma unroll
          for (int j = 0; j < 16; ++j) cw[j] = sl[(size_t)j * 128];
          wait_d(x);
          const bool pe_layer = (l == 3) && (cq == 3);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int cb = cq * 32 + c * 16;
            uint32_t ga[16];
            tc::tmem_ld16(tl + T4_D(x) + cb, ga);
            tc::tmem_wait_ld();
            float a[16];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float d0, d1;
              t4_decode_pair(cw[c * 8 + (j >> 1)], d0, d1);
              a[j] = __uint_as_float(ga[j]) * d0;
              a[j + 1] = __uint_as_float(ga[j + 1]) * d1;
            }
            if (pe_layer) {       // gradient w.r.t. the PE part of the skip layer's input (columns 101..127)
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (96 + c * 16 + j >= 101) {
                  t4_pe_accum(PEs + x * 27 * 128, r, 96 + c * 16 + j - 101, __uint_as_float(ga[j]), g1[x]);
                  a[j] = 0.f;
                }
            }
            store_a(x, cb, a);
          }
          signal(x);
        }
      }
      // ---- lin0's input gradient = PE gradient; feature gradients from F; d sdf / d x; outputs ----
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        wait_d(x);
        float o3[3] = {0.f, 0.f, 0.f};
        {
          uint32_t f0[8], f1[8];
          const int c0 = cq * 7;
          tc::tmem_ld8(tl + T4_F(x) + (c0 & ~7), f0);
          tc::tmem_ld8(tl + T4_F(x) + (c0 & ~7) + 8, f1);
          tc::tmem_wait_ld();
          float gf[7];
          switch (cq) {       // columns 7 cq .. 7 cq + 6 of F
            case 0:
#pragma unroll
              for (int k = 0; k < 7; ++k) gf[k] = __uint_as_float(f0[k]);
              break;
            case 1:
              gf[0] = __uint_as_float(f0[7]);
#pragma unroll
              for (int k = 1; k < 7; ++k) gf[k] = __uint_as_float(f1[k - 1]);
              break;
            case 2:
              gf[0] = __uint_as_float(f0[6]); gf[1] = __uint_as_float(f0[7]);
#pragma unroll
              for (int k = 2; k < 7; ++k) gf[k] = __uint_as_float(f1[k - 2]);
              break;
            default:
              gf[0] = __uint_as_float(f0[5]); gf[1] = __uint_as_float(f0[6]); gf[2] = __uint_as_float(f0[7]);
#pragma unroll
              for (int k = 3; k < 7; ++k) gf[k] = __uint_as_float(f1[k - 3]);
              break;
          }
#pragma unroll
          for (int k = 0; k < 7; ++k) gf[k] = fmaf(net.w6[128 + cq * 7 + k], c6s, gf[k]);     // lin6 sees the features too
          if (P[x].id >= 0 && cq < sc.n_levels) sparse_back_first(sc, cq, P[x].px, P[x].py, P[x].pz, gf, o3);
        }
        if (cq == 0) {
          uint32_t ga[16], gb[16];
          tc::tmem_ld16(tl + T4_D(x), ga);
          tc::tmem_ld16(tl + T4_D(x) + 16, gb);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) t4_pe_accum(PEs + x * 27 * 128, r, j, __uint_as_float(ga[j]), g1[x]);
#pragma unroll
          for (int j = 16; j < 27; ++j) t4_pe_accum(PEs + x * 27 * 128, r, j, __uint_as_float(gb[j - 16]), g1[x]);
        }
        // partial sums of the column quarters -> quarter 0
        float* part = PART + (r * 4 + cq) * 4;
#pragma unroll
        for (int d = 0; d < 3; ++d) part[d] = fmaf(g1[x][d], scale, o3[d]);
        part[3] = head[x];
        tc::tc_fence_before();
        quad_sync();
        if (cq == 0 && P[x].id >= 0) {
          const float* p4 = PART + r * 16;
          float s = (p4[3] + p4[7]) + (p4[11] + p4[15]) + net.b6;
          s *= c6s;
          sdf_out[P[x].id] = negate ? -s : s;
#pragma unroll
          for (int d = 0; d < 3; ++d) grad_out[P[x].id * 3 + d] = (p4[d] + p4[4 + d]) + (p4[8 + d] + p4[12 + d]);
        }
        quad_sync();          // PART is rewritten by the other tile / the next pair
      }
    }
  } else {
    // =============================== weight stream + MMA issue ===============================
    if (tc::elect_one()) {
      const uint32_t w_a = tc::smem_u32(smem + T4_SM_W);
      const uint32_t ab_hi = tc::smem_u32(smem + T4_SM_AB), ab_lo = ab_hi + 32768u;
      const uint32_t ap_hi = tbase + T4_AP_HI, ap_lo = tbase + T4_AP_LO;
      auto load = [&](int step) {
        const uint32_t bytes = st_step_bytes(step);
        tc::mbar_arrive_expect_tx(&bars->w_full, bytes);
        for (uint32_t o = 0; o < bytes; o += 16384u)
          tc::bulk_g2s(smem + T4_SM_W + o, wblob + st_step_off(step) + o, 16384u, &bars->w_full);
      };
      // the MMAs of tile x's step
      auto issue = [&](int x, int step) {
        const uint32_t D = tbase + T4_D(x), F = tbase + T4_F(x);
        const uint32_t af_hi = tc::smem_u32(smem + T4_SM_AF(x)), af_lo = af_hi + 8192u;
        if (step == 0) {                       // lin0 forward: K = 32 (27 PE + ones)
          if (x == 0) st_gemm<128, 2, true>(D, ap_hi, ap_lo, w_a, 8192u, false, fast);
          else st_gemm<128, 2, false>(D, ab_hi, ab_lo, w_a, 8192u, false, fast);
        } else if (step < 6) {                 // lin1..lin5 forward: hidden K = 128 + feature / bias K = 32
          if (x == 0) st_gemm<128, 8, true>(D, ap_hi, ap_lo, w_a, ST_W_HID_LO, false, fast);
          else st_gemm<128, 8, false>(D, ab_hi, ab_lo, w_a, ST_W_HID_LO, false, fast);
          st_gemm<128, 2, false>(D, af_hi, af_lo, w_a + ST_W_FEAT, 8192u, true, fast);
        } else if (step < ST_STEPS - 1) {      // lin5..lin1 reverse: hidden columns -> D, feature columns -> F (+=)
          const bool acc = step != 6;
          if (x == 0) {
            st_gemm<128, 8, true>(D, ap_hi, ap_lo, w_a, ST_W_HID_LO, false, fast);
            st_gemm<32, 8, true>(F, ap_hi, ap_lo, w_a + ST_W_FEAT, 8192u, acc, fast);
          } else {
            st_gemm<128, 8, false>(D, ab_hi, ab_lo, w_a, ST_W_HID_LO, false, fast);
            st_gemm<32, 8, false>(F, ab_hi, ab_lo, w_a + ST_W_FEAT, 8192u, acc, fast);
          }
        } else {                               // lin0 reverse: 27 PE columns
          if (x == 0) st_gemm<32, 8, true>(D, ap_hi, ap_lo, w_a, 8192u, false, fast);
          else st_gemm<32, 8, false>(D, ab_hi, ab_lo, w_a, 8192u, false, fast);
        }
        tc::mma_commit(&bars->d_full[x]);
      };
      uint32_t g = 0;        // steps issued so far = the phase of every barrier
      if ((int64_t)blockIdx.x < n_pairs) load(0);
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll 1
        for (int step = 0; step < ST_STEPS; ++step, ++g) {
          tc::mbar_wait(&bars->a_ready[0], g & 1);
          tc::mbar_wait(&bars->w_full, g & 1);
          tc::tc_fence_after();
          issue(0, step);
          tc::mbar_wait(&bars->a_ready[1], g & 1);
          tc::tc_fence_after();
          issue(1, step);
          // the weight buffer is free once tile B's MMAs are done
          const bool more = step + 1 < ST_STEPS || pair + gridDim.x < n_pairs;
          if (more) {
            tc::mbar_wait(&bars->d_full[1], g & 1);
            load(step + 1 < ST_STEPS ? step + 1 : 0);
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == T4_EPI_WARPS) tc::tmem_dealloc<512>(tbase);
}

int launch_sdf_tc4(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_sdf, float* d_grad,
                   bool negate, bool fast, cudaStream_t st) {
  if (src.n <= 0) return 0;
  SURF_CHECK_ARG(n->smooth_tc_w && n->smooth_tc_scratch, "network without whole-layer tensor-core weights");
  SURF_CHECK_ARG(d_sdf && d_grad, "k_sdf_tc4 writes sdf and gradient");
  int rc = surf_ensure_dyn_smem((const void*)k_sdf_tc4, T4_SM_TOTAL);
  if (rc) return rc;
  const int64_t pairs = ((src.n + ST_ROWS - 1) / ST_ROWS + 1) / 2;
  const int grid = (int)(pairs < n->n_sm ? pairs : n->n_sm);
  surf_time_begin(0, st);
  k_sdf_tc4<<<grid, T4_THREADS, T4_SM_TOTAL, st>>>(s->dev, n->dev, src, n->smooth_tc_w, d_sdf, d_grad,
                                                  reinterpret_cast<uint32_t*>(n->smooth_tc_scratch),
                                                  (negate ? 1 : 0) | (fast ? 2 : 0));
  surf_time_end(0, st);
  SURF_LAUNCH_CHECK();
  return 0;
}
