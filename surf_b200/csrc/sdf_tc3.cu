// K2a + K3, tensor-core edition with TWO 128-point tiles in flight per CTA: SDF MLP forward + analytic input gradient.
//   reference: SDFNetworkSparse.sdf / .gradient (sdf_network.py:95-141), lookup_sparse_volume (projector.py:217-390)
//
// sdf_tc2.cu overlaps the MMAs of layer l+1 with the epilogue of layer l inside ONE tile, which leaves a tail per layer
// (the last column group's MMAs plus two handoffs, ~20 % of the epilogue warps' time) and nothing to do while a tile
// starts or ends.  Two tiles hide both: while the 16 epilogue warps work on tile X, the tensor pipe runs tile Y.
// What makes two tiles fit in the 512 TMEM columns:
//   * the epilogue writes the next A operand IN PLACE over the accumulator columns it has just consumed — a thread
//     reads its 8 fp32 columns and stores [4 columns of fp16 hi | 4 columns of fp16 lo] of the 8 activations there;
//   * such an 8-column block is one K = 16 A operand [hi(k0..7) | lo(k0..7)].  With a B descriptor whose K-group
//     stride (LBO) is 0 both K groups read the SAME 8 weight rows, so A' x dup(W_hi) = hi.W_hi + lo.W_hi and
//     A' x dup(W_lo) = hi.W_lo + lo.W_lo: the full four-term fp16 split product in 2 MMAs per 8 k (the weight stream
//     is unchanged; the MMA count per layer goes from 36 to 38);
//   * a tile therefore occupies ONE 160-column region while its epilogue runs and two (A' in, accumulator out) while
//     its MMAs run; with the two tiles in strict alternation (step n = one layer phase of one tile; X, Y, X, Y ...)
//     three rotating regions suffice: step n accumulates into region n % 3.
// Every layer phase is "all MMAs, then all of the epilogue" for its tile (no column-group signalling).  One elected
// thread issues all MMAs in a fixed order: results are bitwise deterministic.  Helper warps stage a slot's next tile
// (sparse gather + positional encoding -> smem A operands) as soon as the slot's forward pass is over and turn a
// finished tile's feature / PE gradients into d sdf / d x.  Weight stream and softplus' scratch as sdf_tc2.cu.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "surf_internal.cuh"
#include "tc_common.cuh"

#define T3_EPI_WARPS 16
#define T3_EPI_THREADS (T3_EPI_WARPS * 32)
#define T3_HELP_WARPS 4
#define T3_THREADS ((T3_EPI_WARPS + 2 + T3_HELP_WARPS) * 32)     // + 1 MMA issuer + 1 weight loader + helpers
#define T3_SLOT_BYTES 20480                      // N = 160 x K = 32 x (hi + lo)
#define T3_NSLOT 5
#define T3_REGION 160u                           // TMEM columns per region

// dynamic smem (bytes); [2] = per tile slot
#define S3_RING 0
#define S3_AFEAT (S3_RING + T3_NSLOT * T3_SLOT_BYTES)      // [2] x (hi 8 KB | lo 8 KB)  (128 rows x K 32)
#define S3_APE (S3_AFEAT + 2 * 16384)                      // [2]
#define S3_W6 (S3_APE + 2 * 16384)                         // 160 floats
#define S3_PART (S3_W6 + 640)                              // [2][4][128] floats
#define S3_GPE (S3_PART + 2 * 2048)                        // [2][28][128] floats: PE gradient (skip layer + lin0)
#define S3_GF (S3_GPE + 2 * 14336)                         // [2][28][128] floats: feature gradient
#define S3_BAR (S3_GF + 2 * 14336)
#define S3_TOTAL (S3_BAR + 512)

struct T3Bars {
  uint64_t w_full[T3_NSLOT];
  uint64_t w_empty[T3_NSLOT];
  uint64_t d_full[2];       // all MMAs of a layer phase of slot s done (tcgen05.commit)
  uint64_t a_ready[2];      // slot s: the epilogue has written the next A operand: one arrival per epilogue warp
  uint64_t stage_ready[2];  // slot s: smem operands (features, PE) of its next tile staged: one arrival per helper warp
  uint64_t fwd_done[2];     // slot s: forward pass over, its smem operands may be restaged: one arrival per epilogue warp
  uint64_t grads_ready[2];  // slot s: tile through its layers (gradients in smem): one arrival per epilogue warp
  uint64_t finish_done[2];  // slot s: helpers done with the tile's gradients in smem: one arrival per helper warp
  uint32_t tmem_base;
};

__device__ __forceinline__ float t3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float t3_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float t3_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// softplus(beta = 100) and the 16-bit code of e = exp(-|100 z|) for the reverse pass: see sdf_tc2.cu
__device__ __forceinline__ float t3_softplus(float z, float& e) {
  e = t3_ex2(fabsf(z) * -144.26950408889634f);
  return fmaf(t3_lg2(1.0f + e), 0.0069314718055994531f, fmaxf(z, 0.f));
}
__device__ __forceinline__ uint32_t t3_code_word(float e) { return __float_as_uint(fminf(e, 0.9999847412109375f) + 128.0f); }
template <int T>
__device__ __forceinline__ float t3_decode_u(uint32_t pair) {
  return __uint_as_float(__byte_perm(pair, 0x43000000u, T ? 0x7632 : 0x7610)) - 127.0f;
}

__device__ __host__ __forceinline__ int t3_phase_chunks(int phase) {   // 0..5 forward, 6..10 reverse lin5..lin1, 11 reverse lin0
  if (phase == 0) return 1;
  if (phase < 6) return 6;      // 2 x K16 (features | bias) first, then 4 x K32 hidden
  if (phase < 11) return 4;
  return 2;
}

struct T3Epi {
  uint32_t tl;            // TMEM base of my lane quarter
  int part, r, te;
  const uint8_t* ape;
  const float* sw6;
  float inv_scale;
  uint4* scratch;
  uint32_t* sgn_scratch;
  float* s_gpe;
};

__device__ __forceinline__ float t3_get_k(const uint8_t* base, int r, int k) {
  const uint32_t off = (uint32_t)(k >> 3) * 2048u + r * 16 + (k & 7) * 2;
  return __half2float(*reinterpret_cast<const __half*>(base + off)) +
         __half2float(*reinterpret_cast<const __half*>(base + 8192 + off));
}

// Forward activation of my 8 columns cb .. cb+7: h = softplus(z).  SKIP: 0 = plain hidden columns; 1 = columns 5..7
// of my 8 are positional-encoding inputs of the skip layer (cb == 96); 2 = all 8 are.  HEAD: lin5 -> SDF head partial
// sum and (GRAD) h = delta5 = w6 / scale * softplus', the first reverse A operand.
template <bool GRAD, int SKIP, bool HEAD>
__device__ __forceinline__ void t3_fwd_act(const T3Epi& c, int cb, const uint32_t (&d)[8], float (&h)[8], uint4& spw,
                                           uint32_t& sgn, float& head) {
  uint32_t cw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const bool pe = (SKIP == 2) || (SKIP == 1 && n >= 5);
    const float z = __uint_as_float(d[n]);
    if (pe) {
      h[n] = t3_get_k(c.ape, c.r, cb + n - 101);
      if (GRAD) sgn = __funnelshift_l(0x80000000u, sgn, 1);
    } else {
      float e;
      h[n] = t3_softplus(z, e);
      if (HEAD) {
        const float w = c.sw6[cb + n];
        head = fmaf(h[n], w, head);
        if (GRAD) {
          const float rr = t3_rcp(1.0f + e);
          h[n] = w * c.inv_scale * (z >= 0.f ? rr : 1.0f - rr);
        }
      } else if (GRAD) {
        cw[n] = t3_code_word(e);
        sgn = __funnelshift_l(__float_as_uint(z), sgn, 1);      // element e of the layer ends at bit 31 - e
      }
    }
  }
  if (GRAD && !HEAD)
    spw = make_uint4(__byte_perm(cw[0], cw[1], 0x5410), __byte_perm(cw[2], cw[3], 0x5410), __byte_perm(cw[4], cw[5], 0x5410),
                     __byte_perm(cw[6], cw[7], 0x5410));
}

// Reverse activation of my 8 columns: v = delta_{l-1} = D * softplus'(z_{l-1}); the sign of element n is bit 31 - n of
// `sgn`.  SKIP as above: those columns are the PE input gradient of the skip layer (kept in smem, v = 0).
template <int SKIP>
__device__ __forceinline__ void t3_bwd_act(const T3Epi& c, int cb, const uint32_t (&d)[8], const uint4& spw, uint32_t sgn,
                                           float (&v)[8]) {
  const uint32_t sp[4] = {spw.x, spw.y, spw.z, spw.w};
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const bool pe = (SKIP == 2) || (SKIP == 1 && n >= 5);
    const float g = __uint_as_float(d[n]);
    if (pe) {
      c.s_gpe[(cb + n - 101) * 128 + c.r] = g;
      v[n] = 0.f;
    } else {
      const float rr = t3_rcp((n & 1) ? t3_decode_u<1>(sp[n >> 1]) : t3_decode_u<0>(sp[n >> 1]));
      const bool neg = ((sgn >> (31 - n)) & 1u) != 0u;
      v[n] = g * (neg ? 1.0f - rr : rr);
    }
  }
}

// my 8 activations -> the 8 TMEM columns they came from, as one K = 16 A operand [hi(k0..7) | lo(k0..7)]
__device__ __forceinline__ void t3_store_a(uint32_t taddr, const float (&h)[8]) {
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) tc::split2(h[2 * j], h[2 * j + 1], w[j], w[4 + j]);
  tc::tmem_st8(taddr, w);
}

__device__ __forceinline__ void t3_wait_ld(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}

// forward layer l of a tile whose accumulator sits in TMEM region `reg`.  Eight activations per thread and iteration,
// the next group's accumulator columns in flight meanwhile (16 per iteration measured slower: spills at 80 registers
// and the reverse pass loses the one-group-ahead prefetch of the softplus' codes).
template <bool GRAD, bool HEAD>
__device__ __forceinline__ void t3_fwd_layer(const T3Epi& c, uint32_t reg, int l, bool skip_next, float& head) {
  const uint32_t dcol = c.tl + reg * T3_REGION + c.part * 8;
  constexpr bool STORE = !HEAD || GRAD;
  uint32_t sgn = 0;
  uint32_t dn[8];
  tc::tmem_ld8(dcol, dn);
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    t3_wait_ld(dn);
    uint32_t dv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dv[j] = dn[j];
    if (g < 3) tc::tmem_ld8(dcol + (g + 1) * 32, dn);
    const int cb = g * 32 + c.part * 8;
    uint4 spw = make_uint4(0u, 0u, 0u, 0u);
    float h[8];
    if (!HEAD && g == 3 && skip_next) {
      if (c.part == 0) t3_fwd_act<GRAD, 1, false>(c, cb, dv, h, spw, sgn, head);
      else t3_fwd_act<GRAD, 2, false>(c, cb, dv, h, spw, sgn, head);
    } else {
      t3_fwd_act<GRAD, 0, HEAD>(c, cb, dv, h, spw, sgn, head);
    }
    if (STORE) t3_store_a(dcol + g * 32, h);
    if (GRAD && !HEAD) c.scratch[(size_t)(l * 4 + g) * T3_EPI_THREADS + c.te] = spw;
  }
  if (GRAD && !HEAD) c.sgn_scratch[(size_t)l * T3_EPI_THREADS + c.te] = sgn;
}

// reverse layer: region `reg` holds d sdf / d (input of lin_l), columns 128.. the feature-gradient part; lsrc = l - 1
__device__ __forceinline__ void t3_bwd_layer(const T3Epi& c, uint32_t reg, bool skip_pe, int lsrc, float (&gf)[8]) {
  const uint32_t dbase = c.tl + reg * T3_REGION + c.part * 8;
  uint32_t sgn = c.sgn_scratch[(size_t)lsrc * T3_EPI_THREADS + c.te];
  uint4 spw = c.scratch[(size_t)(lsrc * 4) * T3_EPI_THREADS + c.te];
  uint32_t dn[8];
  tc::tmem_ld8(dbase, dn);
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    t3_wait_ld(dn);
    uint32_t dv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dv[j] = dn[j];
    tc::tmem_ld8(dbase + (g + 1) * 32, dn);               // g == 3: the feature-gradient columns 128 + part*8 ..
    const uint4 spw_cur = spw;
    if (g < 3) spw = c.scratch[(size_t)(lsrc * 4 + g + 1) * T3_EPI_THREADS + c.te];     // next group's codes
    const int cb = g * 32 + c.part * 8;
    float v[8];
    if (g == 3 && skip_pe) {
      if (c.part == 0) t3_bwd_act<1>(c, cb, dv, spw_cur, sgn, v);
      else t3_bwd_act<2>(c, cb, dv, spw_cur, sgn, v);
    } else {
      t3_bwd_act<0>(c, cb, dv, spw_cur, sgn, v);
    }
    t3_store_a(dbase + g * 32, v);
    sgn <<= 8;
  }
  t3_wait_ld(dn);
#pragma unroll
  for (int j = 0; j < 8; ++j) gf[j] += __uint_as_float(dn[j]);
}

// The two tile slots advance in strict alternation, one layer phase per step; every role (epilogue, issuer, loader)
// walks the same sequence.  Slot s owns the local tiles s, s + 2, s + 4, ...
struct T3Steps {
  // scalars and selects only: an array indexed by the slot would live in local memory (measured: the stack traffic of
  // 704 threads thrashes the little L1 the 225 KB of shared memory leave and ends up in DRAM)
  int64_t n0, n1, j0, j1;
  int ph0, ph1;
  int s;            // slot of the current step
  int64_t n;        // step counter: the step accumulates into TMEM region n % 3
  int nphase;
  __device__ __forceinline__ void init(int64_t my_tiles, int nphase_) {
    n0 = (my_tiles + 1) >> 1; n1 = my_tiles >> 1;
    j0 = j1 = 0; ph0 = ph1 = 0; s = 1; n = -1; nphase = nphase_;
  }
  __device__ __forceinline__ int phase() const { return s ? ph1 : ph0; }
  __device__ __forceinline__ int64_t tile() const { return s ? j1 : j0; }
  // advance to the next step; false when both slots are out of tiles
  __device__ __forceinline__ bool next() {
    if (n >= 0) {                       // retire the step just done
      if (s) { if (++ph1 == nphase) { ph1 = 0; j1++; } }
      else   { if (++ph0 == nphase) { ph0 = 0; j0++; } }
    }
    const bool other_has = s ? (j0 < n0) : (j1 < n1);
    const bool mine_has = s ? (j1 < n1) : (j0 < n0);
    if (other_has) s ^= 1;
    else if (!mine_has) return false;
    n++;
    return true;
  }
};

template <bool GRAD>
__global__ void __launch_bounds__(T3_THREADS, 1)
k_sdf_tc3(const DevScene sc, const DevNet net, const PointSource src, const uint8_t* __restrict__ wblob,
          const T1Stream stream, float* __restrict__ sdf_out, float* __restrict__ grad_out,
          uint4* __restrict__ scratch_all, int flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  T3Bars* bars = reinterpret_cast<T3Bars*>(smem + S3_BAR);
  float* sw6 = reinterpret_cast<float*>(smem + S3_W6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NPHASE = GRAD ? 12 : 6;

  int64_t n_total = src.n;
  if (src.count) {
    const int64_t c = *src.count;
    n_total = c < n_total ? c : n_total;
  }
  const int64_t n_tiles = (n_total + 127) / 128;
  int64_t my_tiles = 0;
  if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == T3_EPI_WARPS) tc::tmem_alloc<512>(&bars->tmem_base);
  if (tid == 0) {
    for (int i = 0; i < T3_NSLOT; ++i) {
      tc::mbar_init(&bars->w_full[i], 1);
      tc::mbar_init(&bars->w_empty[i], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bars->d_full[s], 1);
      tc::mbar_init(&bars->a_ready[s], T3_EPI_WARPS);
      tc::mbar_init(&bars->stage_ready[s], T3_HELP_WARPS);
      tc::mbar_init(&bars->fwd_done[s], T3_EPI_WARPS);
      tc::mbar_init(&bars->grads_ready[s], T3_EPI_WARPS);
      tc::mbar_init(&bars->finish_done[s], T3_HELP_WARPS);
    }
    tc::mbar_fence_init();
  }
  for (int i = tid; i < 160; i += T3_THREADS) sw6[i] = net.w6[i];
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = bars->tmem_base;

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
  // local tile index of slot s's j-th tile -> global tile
  auto tile_of = [&](int s, int64_t j) -> int64_t { return (int64_t)blockIdx.x + (2 * j + s) * (int64_t)gridDim.x; };

  if (warp < T3_EPI_WARPS) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, part = warp >> 2;
    const int r = q * 32 + lane;                 // row = point = TMEM lane
    const uint32_t tl = tbase + ((uint32_t)(q * 32) << 16);
    const int te = warp * 32 + lane;             // 0..511
    const size_t scratch_slot = (size_t)(5 * 4 * T3_EPI_THREADS + 5 * T3_EPI_THREADS / 4);   // uint4 per slot
    uint4* scratch0 = scratch_all + (size_t)blockIdx.x * 2 * scratch_slot;
    auto epi_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory"); };
    auto warp_arrive = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar);
    };

    T3Epi ec;
    ec.tl = tl; ec.part = part; ec.r = r; ec.te = te; ec.sw6 = sw6; ec.inv_scale = net.inv_scale;
    float gfa[8], gfb[8];                         // reverse pass, per slot: d sdf / d feat of columns part*8 .. +7
#pragma unroll
    for (int j = 0; j < 8; ++j) gfa[j] = gfb[j] = 0.f;
    uint32_t ph_d0 = 0u, ph_d1 = 0u;

    T3Steps st;
    st.init(my_tiles, NPHASE);
    while (st.next()) {
      const int s = st.s, p = st.phase();
      const int64_t j = st.tile();
      const uint32_t reg = (uint32_t)(st.n % 3);
      const uint8_t* afeat = smem + S3_AFEAT + s * 16384;
      float* s_part = reinterpret_cast<float*>(smem + S3_PART) + s * 512;
      float* s_gpe = reinterpret_cast<float*>(smem + S3_GPE) + s * 3584;
      float* s_gf = reinterpret_cast<float*>(smem + S3_GF) + s * 3584;
      ec.ape = smem + S3_APE + s * 16384;
      ec.scratch = scratch0 + s * scratch_slot;
      ec.sgn_scratch = reinterpret_cast<uint32_t*>(ec.scratch + 5 * 4 * T3_EPI_THREADS);
      ec.s_gpe = s_gpe;
      if (p == 0) {
        // the helpers staged this tile long ago; the wait is the acquire for my reads of their smem writes
        tc::mbar_wait(&bars->stage_ready[s], (uint32_t)(j & 1));
      }
      if (GRAD && p == 6 && j > 0) {
        // the helpers must be done reading the slot's previous gradients before this tile overwrites them
        tc::mbar_wait(&bars->finish_done[s], (uint32_t)((j - 1) & 1));
      }
      tc::mbar_wait(&bars->d_full[s], (s ? ph_d1 : ph_d0) & 1);
      if (s) ph_d1++; else ph_d0++;
      tc::tc_fence_after();

      if (p < 6) {
        // ------------------------------------ forward layer p ------------------------------------
        float head = 0.f;
        if (p < 5) t3_fwd_layer<GRAD, false>(ec, reg, p, p + 1 == net.skip_layer, head);
        else t3_fwd_layer<GRAD, true>(ec, reg, p, false, head);
        if (p < 5 || GRAD) {
          tc::tmem_wait_st();
          tc::tc_fence_before();
          warp_arrive(&bars->a_ready[s]);
        }
        if (p == 5) {
          s_part[part * 128 + r] = head;
          epi_bar();
          if (part == 0) {
            float sv = s_part[r] + s_part[128 + r] + s_part[256 + r] + s_part[384 + r] + net.b6;
#pragma unroll
            for (int c = 0; c < 28; ++c) sv = fmaf(t3_get_k(afeat, r, c), sw6[128 + c], sv);
            sv *= net.inv_scale;
            const int64_t i = tile_of(s, j) * 128 + r;
            if (i < n_total) {
              const int64_t id = src.list ? (int64_t)src.list[i] : i;
              sdf_out[id] = (flags & 1) ? -sv : sv;
            }
          }
          // forward pass over: the slot's smem operands may be restaged for its next tile
          tc::tc_fence_before();
          warp_arrive(&bars->fwd_done[s]);
          if (!GRAD) {
            if (j > 0) tc::mbar_wait(&bars->finish_done[s], (uint32_t)((j - 1) & 1));
            warp_arrive(&bars->grads_ready[s]);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float g0 = sw6[128 + part * 8 + k] * net.inv_scale;
              if (s == 0) gfa[k] = g0; else gfb[k] = g0;
            }
          }
        }
      } else if (p < 11) {
        // ------------------------------------ reverse of lin_l, l = 11 - p ------------------------------------
        const int l = 11 - p;
        if (s == 0) t3_bwd_layer(ec, reg, l == net.skip_layer, l - 1, gfa);
        else t3_bwd_layer(ec, reg, l == net.skip_layer, l - 1, gfb);
        tc::tmem_wait_st();
        tc::tc_fence_before();
        warp_arrive(&bars->a_ready[s]);
        if (p == 10) {
          // feature gradients -> smem (28 x 128)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (part * 8 + k < 28) s_gf[(part * 8 + k) * 128 + r] = s == 0 ? gfa[k] : gfb[k];
        }
      } else {
        // ---- reverse of lin0: PE gradient through lin0 = delta0 . W0 (N = 32), added to the skip layer's ----
        epi_bar();                       // every warp's skip-layer s_gpe rows are written
        if (part == 0) {
          uint32_t a[16];
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            tc::tmem_ld16(tl + reg * T3_REGION + hb * 16, a);
            tc::tmem_wait_ld();
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
              const int k = hb * 16 + k2;
              if (k < 27) s_gpe[k * 128 + r] += __uint_as_float(a[k2]);
            }
          }
        }
        tc::tc_fence_before();
        warp_arrive(&bars->grads_ready[s]);       // the helpers turn s_gf / s_gpe into d sdf / d x
      }
    }
  } else if (warp == T3_EPI_WARPS) {
    // =============================== the MMA issuer ===============================
    const bool fast = (flags & 2) != 0;       // only the hi-weight MMA of every block (opt-in reduced-precision mode)
    if (tc::elect_one()) {
      const uint32_t ring = tc::smem_u32(smem + S3_RING);
      const uint32_t id128 = tc::idesc_f16(128, 128, 0), id160 = tc::idesc_f16(128, 160, 0), id32 = tc::idesc_f16(128, 32, 0);
      // B descriptors: SBO 128; LBO = rows * 16 for the ordinary K = 16 step, LBO = 0 for the duplicated 8-row K group
      const uint64_t d128 = tc::smem_desc_kmajor(0, 2048, 128), dup = tc::smem_desc_kmajor(0, 0, 128);
      const uint32_t dh128 = (uint32_t)(d128 >> 32), dhdup = (uint32_t)(dup >> 32);
      const uint32_t afeat_lo0 = (uint32_t)d128 | (tc::smem_u32(smem + S3_AFEAT) >> 4);
      const uint32_t ape_lo0 = (uint32_t)d128 | (tc::smem_u32(smem + S3_APE) >> 4);
      uint32_t ph_a0 = 0u, ph_a1 = 0u;
      int slot = 0;
      uint32_t ring_par = 0;
      auto next_chunk = [&]() -> uint32_t {
        tc::mbar_wait(&bars->w_full[slot], ring_par);
        return (ring + slot * T3_SLOT_BYTES) >> 4;
      };
      auto release_chunk = [&]() {
        tc::mma_commit(&bars->w_empty[slot]);
        slot = (slot + 1 == T3_NSLOT) ? 0 : slot + 1;
        ring_par ^= (slot == 0);
      };
      uint32_t reg_in0 = 0u, reg_in1 = 0u;      // region holding slot s's current A operand
      T3Steps st;
      st.init(my_tiles, NPHASE);
      while (st.next()) {
        const int s = st.s, p = st.phase();
        const int64_t j = st.tile();
        const uint32_t reg = (uint32_t)(st.n % 3);
        const uint32_t tD = tbase + reg * T3_REGION;
        const uint32_t tA = tbase + (s ? reg_in1 : reg_in0) * T3_REGION;
        if (p == 0) {
          tc::mbar_wait(&bars->stage_ready[s], (uint32_t)(j & 1));
          // every epilogue warp must have consumed the slot's previous last d_full phase before it completes again
          if (j > 0) tc::mbar_wait(&bars->grads_ready[s], (uint32_t)((j - 1) & 1));
          tc::tc_fence_after();
          // ---- lin0 on the positional encoding (A in smem) ----
          const uint32_t w0 = (uint32_t)d128 | next_chunk(), dh = dh128, a0 = ape_lo0 + s * 1024;
          if (!fast) {
            tc::mma_ss_w<false>(tD, a0 + 512, dh, w0, dh, id128);
            tc::mma_ss_w<true>(tD, a0, dh, w0 + 512, dh, id128);
            tc::mma_ss_w<true>(tD, a0 + 768, dh, w0 + 256, dh, id128);
            tc::mma_ss_w<true>(tD, a0 + 256, dh, w0 + 768, dh, id128);
            tc::mma_ss_w<true>(tD, a0, dh, w0, dh, id128);
          } else {
            tc::mma_ss_w<false>(tD, a0, dh, w0, dh, id128);
          }
          tc::mma_ss_w<true>(tD, a0 + 256, dh, w0 + 256, dh, id128);
          release_chunk();
        } else {
          tc::mbar_wait(&bars->a_ready[s], (s ? ph_a1 : ph_a0) & 1);
          if (s) ph_a1++; else ph_a0++;
          tc::tc_fence_after();
          if (p < 6) {
            // feature | bias columns (A in smem, fp16 hi / lo halves): two K = 16 half chunks
            const uint32_t dh = dh128, afeat_lo = afeat_lo0 + s * 1024;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t w0 = (uint32_t)d128 | next_chunk();
              const uint32_t a0 = afeat_lo + h * 256;
              if (!fast) {
                if (h == 0) tc::mma_ss_w<false>(tD, a0 + 512, dh, w0, dh, id128); else tc::mma_ss_w<true>(tD, a0 + 512, dh, w0, dh, id128);
                tc::mma_ss_w<true>(tD, a0, dh, w0 + 256, dh, id128);
                tc::mma_ss_w<true>(tD, a0, dh, w0, dh, id128);
              } else {
                if (h == 0) tc::mma_ss_w<false>(tD, a0, dh, w0, dh, id128); else tc::mma_ss_w<true>(tD, a0, dh, w0, dh, id128);
              }
              release_chunk();
            }
            // hidden columns: chunk c = 4 blocks of 8 k; block = A' [hi | lo] x dup(W_lo) then x dup(W_hi)
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              const uint32_t wa = next_chunk();                    // hi half at +0, lo half at +8192 B (512 units)
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const uint32_t a = tA + c * 32 + b * 8;
                const uint32_t wb = wa + b * 128;                  // K group b: 128 rows x 16 B
                if (!fast) tc::mma_ts_w<true>(tD, a, wb + 512, dhdup, id128);
                tc::mma_ts_w<true>(tD, a, wb, dhdup, id128);
              }
              release_chunk();
            }
          } else if (p < 11) {
            // reverse of lin5..lin1: N = 160; lo half at +10240 B (640 units), K group = 160 rows x 16 B (160 units)
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              const uint32_t wa = next_chunk();
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const uint32_t a = tA + c * 32 + b * 8;
                const uint32_t wb = wa + b * 160;
                if (!fast) {
                  if (c == 0 && b == 0) tc::mma_ts_w<false>(tD, a, wb + 640, dhdup, id160); else tc::mma_ts_w<true>(tD, a, wb + 640, dhdup, id160);
                  tc::mma_ts_w<true>(tD, a, wb, dhdup, id160);
                } else {
                  if (c == 0 && b == 0) tc::mma_ts_w<false>(tD, a, wb, dhdup, id160); else tc::mma_ts_w<true>(tD, a, wb, dhdup, id160);
                }
              }
              release_chunk();
            }
          } else {
            // reverse of lin0: N = 32, two K = 64 chunks; lo half at +4096 B (256 units), K group = 32 rows x 16 B
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              const uint32_t wa = next_chunk();
#pragma unroll
              for (int b = 0; b < 8; ++b) {
                const uint32_t a = tA + c * 64 + b * 8;
                const uint32_t wb = wa + b * 32;
                if (!fast) {
                  if (c == 0 && b == 0) tc::mma_ts_w<false>(tD, a, wb + 256, dhdup, id32); else tc::mma_ts_w<true>(tD, a, wb + 256, dhdup, id32);
                  tc::mma_ts_w<true>(tD, a, wb, dhdup, id32);
                } else {
                  if (c == 0 && b == 0) tc::mma_ts_w<false>(tD, a, wb, dhdup, id32); else tc::mma_ts_w<true>(tD, a, wb, dhdup, id32);
                }
              }
              release_chunk();
            }
          }
        }
        tc::mma_commit(&bars->d_full[s]);
        if (s) reg_in1 = reg; else reg_in0 = reg;
      }
    }
  } else if (warp >= T3_EPI_WARPS + 2) {
    // =============================== helper warps: staging and the final gradient ===============================
    const int hl = (warp - (T3_EPI_WARPS + 2)) * 32 + lane;       // one point (row) per helper thread
    auto put_k = [&](uint8_t* base, int r, int k, float v) {
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const uint32_t off = (uint32_t)(k >> 3) * 2048u + r * 16 + (k & 7) * 2;
      *reinterpret_cast<__half*>(base + off) = h;
      *reinterpret_cast<__half*>(base + 8192 + off) = l;
    };
    auto warp_arrive = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar);
    };
    // features (4 levels x 7, trilinear) and positional encoding of slot s's j-th tile -> the slot's smem A operands
    auto stage = [&](int s, int64_t j) {
      uint8_t* afeat = smem + S3_AFEAT + s * 16384;
      uint8_t* ape = smem + S3_APE + s * 16384;
      const int r = hl;
      float px, py, pz;
      load_point(tile_of(s, j) * 128 + r, px, py, pz);
#pragma unroll
      for (int lv = 0; lv < 4; ++lv) {
        float f7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (lv < sc.n_levels) sparse_level<0>(sc, lv, px, py, pz, nullptr, f7);
#pragma unroll
        for (int c = 0; c < 7; ++c) put_k(afeat, r, lv * 7 + c, f7[c]);
      }
      put_k(afeat, r, 28, 1.0f);
#pragma unroll
      for (int k = 29; k < 32; ++k) put_k(afeat, r, k, 0.f);
      const float xs[3] = {px * net.scale, py * net.scale, pz * net.scale};
#pragma unroll
      for (int d = 0; d < 3; ++d) put_k(ape, r, d, xs[d]);
      float fr = 1.0f;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          float sn = 0.f, cs = 0.f;
          if (f < net.multires) sincosf(xs[d] * fr, &sn, &cs);
          put_k(ape, r, 3 + 6 * f + d, sn);
          put_k(ape, r, 3 + 6 * f + 3 + d, cs);
        }
        fr *= 2.0f;
      }
      put_k(ape, r, 27, 1.0f);
#pragma unroll
      for (int k = 28; k < 32; ++k) put_k(ape, r, k, 0.f);
      tc::fence_proxy_async();
      warp_arrive(&bars->stage_ready[s]);
    };
    // d sdf / d x of slot s's j-th tile = scale * (d PE / d x)^T g_pe + sum over levels (d feat / d x)^T g_feat
    auto finish = [&](int s, int64_t j) {
      const float* s_gpe = reinterpret_cast<const float*>(smem + S3_GPE) + s * 3584;
      const float* s_gf = reinterpret_cast<const float*>(smem + S3_GF) + s * 3584;
      const int r = hl;
      float px, py, pz;
      const int64_t id = load_point(tile_of(s, j) * 128 + r, px, py, pz);
      float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int lv = 0; lv < 4; ++lv) {
        if (lv < sc.n_levels) {
          float g7[7], o3[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 7; ++c) g7[c] = s_gf[(lv * 7 + c) * 128 + r];
          sparse_level<1>(sc, lv, px, py, pz, g7, o3);
          acc[0] += o3[0]; acc[1] += o3[1]; acc[2] += o3[2];
        }
      }
      const float xs[3] = {px * net.scale, py * net.scale, pz * net.scale};
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float gx = s_gpe[d * 128 + r];
        float fr = 1.0f;
        for (int f = 0; f < net.multires; ++f) {
          // the same fp16 hi + lo rounding of sin / cos as the forward operand saw
          float sn, cs;
          sincosf(xs[d] * fr, &sn, &cs);
          const __half sh = __float2half_rn(sn), ch = __float2half_rn(cs);
          sn = __half2float(sh) + __half2float(__float2half_rn(sn - __half2float(sh)));
          cs = __half2float(ch) + __half2float(__float2half_rn(cs - __half2float(ch)));
          gx += fr * (s_gpe[(3 + 6 * f + d) * 128 + r] * cs - s_gpe[(3 + 6 * f + 3 + d) * 128 + r] * sn);
          fr *= 2.0f;
        }
        gx = fmaf(gx, net.scale, acc[d]);
        if (id >= 0) grad_out[id * 3 + d] = gx;
      }
    };
    const int64_t n0 = (my_tiles + 1) >> 1, n1 = my_tiles >> 1;
    if (n0 > 0) stage(0, 0);
    if (n1 > 0) stage(1, 0);
    for (int64_t j = 0; j < n0; ++j) {
      // events in time order: forward pass of (0, j) over, of (1, j) over, then (0, j) finished, (1, j) finished
      if (j + 1 < n0) {
        tc::mbar_wait(&bars->fwd_done[0], (uint32_t)(j & 1));
        stage(0, j + 1);
      }
      if (j + 1 < n1) {
        tc::mbar_wait(&bars->fwd_done[1], (uint32_t)(j & 1));
        stage(1, j + 1);
      }
      tc::mbar_wait(&bars->grads_ready[0], (uint32_t)(j & 1));
      if (GRAD) finish(0, j);
      warp_arrive(&bars->finish_done[0]);
      if (j < n1) {
        tc::mbar_wait(&bars->grads_ready[1], (uint32_t)(j & 1));
        if (GRAD) finish(1, j);
        warp_arrive(&bars->finish_done[1]);
      }
    }
  } else {
    // =============================== weight loader ===============================
    if (lane == 0) {
      int slot = 0;
      uint32_t par = 1;          // parity of the previous use of this slot (first round: nothing to wait for)
      int64_t loaded = 0;
      T3Steps st;
      st.init(my_tiles, NPHASE);
      while (st.next()) {
        const int p = st.phase();
        int cid = 0;
        for (int q = 0; q < p; ++q) cid += t3_phase_chunks(q);
        const int nch = t3_phase_chunks(p);
        for (int c = 0; c < nch; ++c, ++cid, ++loaded) {
          if (loaded >= T3_NSLOT) tc::mbar_wait(&bars->w_empty[slot], par);
          tc::mbar_arrive_expect_tx(&bars->w_full[slot], stream.bytes[cid]);
          tc::bulk_g2s(smem + S3_RING + slot * T3_SLOT_BYTES, wblob + stream.off[cid], stream.bytes[cid],
                       &bars->w_full[slot]);
          slot = (slot + 1 == T3_NSLOT) ? 0 : slot + 1;
          par ^= (slot == 0);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == T3_EPI_WARPS) tc::tmem_dealloc<512>(tbase);
}

// the weight stream of sdf_tc1.cu with the feature | bias half chunks of every forward layer moved to the front
static T1Stream t3_reordered_stream() {
  T1Stream S = g_t1_stream;
  for (int l = 1; l < 6; ++l) {
    const int base = 1 + (l - 1) * 6;
    const int order[6] = {4, 5, 0, 1, 2, 3};
    for (int c = 0; c < 6; ++c) {
      S.off[base + c] = g_t1_stream.off[base + order[c]];
      S.bytes[base + c] = g_t1_stream.bytes[base + order[c]];
    }
  }
  return S;
}

int launch_sdf_tc3(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_sdf, float* d_grad,
                   bool negate, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SURF_CUDA(cudaFuncSetAttribute(k_sdf_tc3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S3_TOTAL));
    SURF_CUDA(cudaFuncSetAttribute(k_sdf_tc3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S3_TOTAL));
    attr_set = true;
  }
  if (src.n <= 0) return 0;
  const T1Stream stream = t3_reordered_stream();
  const int64_t tiles = (src.n + 127) / 128;
  const int grid = (int)(tiles < n->n_sm ? tiles : n->n_sm);
  const int flags = (negate ? 1 : 0) | (surf_mlp_mode() == 4 ? 2 : 0);
  surf_time_begin(d_grad ? 0 : 1, st);
  if (d_grad) {
    k_sdf_tc3<true><<<grid, T3_THREADS, S3_TOTAL, st>>>(s->dev, n->dev, src, n->tc1_blob, stream, d_sdf, d_grad,
                                                        (uint4*)n->tc1_scratch, flags);
  } else {
    k_sdf_tc3<false><<<grid, T3_THREADS, S3_TOTAL, st>>>(s->dev, n->dev, src, n->tc1_blob, stream, d_sdf, nullptr,
                                                         (uint4*)n->tc1_scratch, flags);
  }
  surf_time_end(d_grad ? 0 : 1, st);
  SURF_LAUNCH_CHECK();
  return 0;
}
