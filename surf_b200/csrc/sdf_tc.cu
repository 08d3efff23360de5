// K2a + K3, tensor-core edition (tcgen05 / TMEM), forward pass: sparse trilinear gather + SDF MLP.
//   reference: SDFNetworkSparse.sdf (sdf_network.py:95-124), lookup_sparse_volume (projector.py:217-390).
//
// One persistent CTA per SM runs TWO independent 128-point tile pipelines (X, Y), each with its own epilogue
// warps, MMA issuer, weight loader and ring; the tensor core interleaves them.  Per tile the fp32 accumulator of a
// layer lives in 128 TMEM columns; the epilogue warps (one thread per point = one TMEM lane) read it with
// tcgen05.ld, apply softplus, split the activation into fp16 hi + lo and write it straight back to TMEM as
// the next layer's A operand (2 x 64 columns) — activations never touch shared memory.  A tile therefore
// owns 256 TMEM columns and the pair fills the 512.  The 28 feature columns + the bias column (K = 32) and the
// layer-0 positional encoding are small shared-memory A operands (SS form).  Weights stream once per layer
// from L2 through an 8-slot ring of 16 KB chunks (cp.async.bulk -> mbarrier) and are used by both tiles; the MMA
// issue order X(l) Y(l) X(l+1) ... overlaps each tile's epilogue with the other tile's MMAs.
// fp32-grade accuracy: every product is formed as hi*hi + lo*hi + hi*lo (3 fp16 MMAs, fp32 accumulate).
#include <math.h>
#include <string.h>

#include <vector>

#include "surf_internal.cuh"
#include "tc_common.cuh"

#define TC_EPI_WARPS 16
#define TC_THREADS ((TC_EPI_WARPS + 5) * 32)   // + 4 MMA issuers (two per tile) + 1 weight loader
#define TC_CHUNK_BYTES 16384          // hi 8 KB + lo 8 KB : K = 32 rows x N = 128
#define TC_HALF_BYTES 8192
#define TC_NSLOT 8                    // ring slots (shared by the two tiles: each chunk is loaded once, used twice)
#define TC_LBO 2048u                  // 128 rows x 16 B : next 8-wide K group
#define TC_SBO 128u
#define TC_CHUNKS_FWD 26              // 1 (lin0) + 5 x 5 (lin1..lin5)

// dynamic smem layout (bytes)
#define TS_RING 0
#define TS_AFEAT (TS_RING + TC_NSLOT * TC_CHUNK_BYTES)     // [2 tiles][hi 8 KB | lo 8 KB]
#define TS_APE (TS_AFEAT + 2 * TC_CHUNK_BYTES)
#define TS_W6 (TS_APE + 2 * TC_CHUNK_BYTES)                // 160 floats
#define TS_PART (TS_W6 + 160 * 4)                          // [2][128] floats: head partial sums
#define TS_BAR (TS_PART + 2 * 128 * 4)                     // barriers
#define TS_TOTAL (TS_BAR + 256)

// optional timeline trace of CTA 0 (debug builds: -DTC_TRACE): (event id, clock) pairs
#ifdef TC_TRACE
__device__ long long g_tc_trace[4096];
__device__ int g_tc_trace_n;
#define TRACE(ev)                                                                  \
  do {                                                                             \
    if (blockIdx.x == 0 && lane == 0) {                                            \
      const int _i = atomicAdd(&g_tc_trace_n, 1);                                  \
      if (_i < 2048) { g_tc_trace[2 * _i] = (ev); g_tc_trace[2 * _i + 1] = clock64(); } \
    }                                                                              \
  } while (0)
#else
#define TRACE(ev) do {} while (0)
#endif

struct TcBars {
  uint64_t w_full[TC_NSLOT];
  uint64_t w_empty[TC_NSLOT];
  uint64_t d_full[2];
  uint64_t a_ready[2];
  uint64_t f_done[2];        // per layer: the accumulate=false MMA of issuer A has completed -> issuer B may start
  uint64_t stagger;          // one-shot: tile Y's issuer starts half a layer period after tile X's
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// softplus(beta = 100) = max(z,0) + log1p(exp(-|100 z|)) / 100, branch-free.  torch's threshold (100 z > 20 -> z,
// sdf_network.py:93) needs no branch: there exp(-100 z) < 2.1e-9, 1 + e rounds to 1 and the log term is exactly 0.
__device__ __forceinline__ float softplus100_fwd(float z) {
  const float e = ex2_ftz(fabsf(z) * -144.26950408889634f);          // exp(-|100 z|)
  return fmaf(lg2_ftz(1.0f + e), 0.0069314718055994531f, fmaxf(z, 0.f));
}

// write 32 fp32 values of one row as fp16 hi/lo into a K-major canonical smem operand (row r, k = 0..31)
__device__ __forceinline__ void store_row_k32(uint8_t* base_hi, int r, const float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc::split2(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1], hi[j], lo[j]);
    *reinterpret_cast<uint4*>(base_hi + g * TC_LBO + r * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base_hi + TC_HALF_BYTES + g * TC_LBO + r * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_sdf_tc_fwd(const DevScene sc, const DevNet net, const PointSource src, const uint8_t* __restrict__ wblob,
             float* __restrict__ sdf_out, int negate, int two_issuers) {
  extern __shared__ __align__(1024) uint8_t smem[];
  TcBars* bars = reinterpret_cast<TcBars*>(smem + TS_BAR);
  float* sw6 = reinterpret_cast<float*>(smem + TS_W6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int64_t n_total = src.n;
  if (src.count) {
    const int64_t c = *src.count;
    n_total = c < n_total ? c : n_total;
  }
  const int64_t n_tiles = (n_total + 127) / 128;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  int64_t my_pairs = 0;
  if ((int64_t)blockIdx.x < n_pairs) my_pairs = (n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == TC_EPI_WARPS) tc::tmem_alloc<512>(&bars->tmem_base);
  if (tid == 0) {
    for (int i = 0; i < TC_NSLOT; ++i) {
      tc::mbar_init(&bars->w_full[i], 1);
      tc::mbar_init(&bars->w_empty[i], 2);       // released by both tiles' MMA issuers
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars->d_full[i], 2);          // both issuers of the tile
      tc::mbar_init(&bars->f_done[i], 1);
      tc::mbar_init(&bars->a_ready[i], 256);
    }
    tc::mbar_init(&bars->stagger, 1);
    tc::mbar_fence_init();
  }
  for (int i = tid; i < 160; i += TC_THREADS) sw6[i] = net.w6[i];
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = bars->tmem_base;

  if (warp < TC_EPI_WARPS) {
    // ============================ epilogue / staging group of one tile ============================
    // 8 warps per tile: warp pair (q, q+4) shares TMEM lane quarter q; `half` selects 64 of the 128 columns.
    const int g = warp >> 3;                    // 0: tile X, 1: tile Y
    const int q = warp & 3, half = (warp >> 2) & 1;
    const int r = q * 32 + lane;                // row = point = TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t tD = tbase + lane_base + g * 256;
    const uint32_t tAhi = tD + 128, tAlo = tD + 192;
    uint8_t* afeat = smem + TS_AFEAT + g * TC_CHUNK_BYTES;
    uint8_t* ape = smem + TS_APE + g * TC_CHUNK_BYTES;
    float* part = reinterpret_cast<float*>(smem + TS_PART) + g * 128;
    uint32_t ph_d = 0;
    auto put_k = [&](uint8_t* base, int k, float v) {       // one element of a K-major canonical smem operand
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const uint32_t off = (uint32_t)(k >> 3) * TC_LBO + r * 16 + (k & 7) * 2;
      *reinterpret_cast<__half*>(base + off) = h;
      *reinterpret_cast<__half*>(base + TC_HALF_BYTES + off) = l;
    };
    auto get_k = [&](const uint8_t* base, int k) {
      const uint32_t off = (uint32_t)(k >> 3) * TC_LBO + r * 16 + (k & 7) * 2;
      return __half2float(*reinterpret_cast<const __half*>(base + off)) +
             __half2float(*reinterpret_cast<const __half*>(base + TC_HALF_BYTES + off));
    };
    for (int64_t it = 0; it < my_pairs; ++it) {
      const int64_t tile = 2 * ((int64_t)blockIdx.x + it * gridDim.x) + g;
      const int64_t i = tile * 128 + r;
      float px = 0.f, py = 0.f, pz = 0.f;
      int64_t id = -1;
      if (i < n_total) {
        id = src.list ? (int64_t)src.list[i] : i;
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
      }
      // ---- stage the two shared-memory A operands: [feats28, 1, 0,0,0] and [PE27, 1, 0,0,0,0];
      //      the two threads of a row split the work: levels {0,1} / {2,3}, PE frequencies {0,1} / {2,3}
#pragma unroll
      for (int ll = 0; ll < 2; ++ll) {
        const int l = half * 2 + ll;
        float f7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (l < sc.n_levels) sparse_level<0>(sc, l, px, py, pz, nullptr, f7);
#pragma unroll
        for (int c = 0; c < 7; ++c) put_k(afeat, l * 7 + c, f7[c]);
      }
      {
        const float xs[3] = {px * net.scale, py * net.scale, pz * net.scale};
        if (half == 0) {
#pragma unroll
          for (int d = 0; d < 3; ++d) put_k(ape, d, xs[d]);
        } else {
          put_k(afeat, 28, 1.0f);
          put_k(ape, 27, 1.0f);
#pragma unroll
          for (int k = 29; k < 32; ++k) put_k(afeat, k, 0.f);
#pragma unroll
          for (int k = 28; k < 32; ++k) put_k(ape, k, 0.f);
        }
#pragma unroll
        for (int ff = 0; ff < 2; ++ff) {
          const int f = half * 2 + ff;
          const float fr = (float)(1 << f);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float sn = 0.f, cs = 0.f;
            if (f < net.multires) sincosf(xs[d] * fr, &sn, &cs);
            put_k(ape, 3 + 6 * f + d, sn);
            put_k(ape, 3 + 6 * f + 3 + d, cs);
          }
        }
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(&bars->a_ready[g]);
      if ((warp & 7) == 0) TRACE(100 * g + 1);          // staging done
      // ---- layers ----
      for (int l = 0; l < 6; ++l) {
        tc::mbar_wait(&bars->d_full[g], ph_d & 1);
        ph_d++;
        tc::tc_fence_after();
        if ((warp & 7) == 0) TRACE(100 * g + 10 + l);     // d_full observed for layer l
        if (l < 5) {
          const bool to_skip = (l + 1 == net.skip_layer);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int cb = half * 2 + cc;       // 32-column block
            uint32_t acc[32];
            tc::tmem_ld32(tD + cb * 32, acc);
            tc::tmem_wait_ld();
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float h0 = softplus100_fwd(__uint_as_float(acc[2 * j]));
              float h1 = softplus100_fwd(__uint_as_float(acc[2 * j + 1]));
              if (to_skip && cb == 3) {   // columns 101..127 of the skip layer's input are the positional encoding
                const int n0 = 96 + 2 * j;
                if (n0 >= 101) h0 = get_k(ape, n0 - 101);
                if (n0 + 1 >= 101) h1 = get_k(ape, n0 + 1 - 101);
              }
              tc::split2(h0, h1, hi[j], lo[j]);
            }
            tc::tmem_st16(tAhi + cb * 16, hi);
            tc::tmem_st16(tAlo + cb * 16, lo);
          }
          tc::tmem_wait_st();
          tc::tc_fence_before();
          tc::mbar_arrive(&bars->a_ready[g]);
          if ((warp & 7) == 0) TRACE(100 * g + 20 + l);   // epilogue of layer l done
        } else {
          float s = 0.f;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int cb = half * 2 + cc;
            uint32_t acc[32];
            tc::tmem_ld32(tD + cb * 32, acc);
            tc::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) s = fmaf(softplus100_fwd(__uint_as_float(acc[j])), sw6[cb * 32 + j], s);
          }
          if (half == 1) part[r] = s;
          tc::tc_fence_before();
          asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");    // the 8 warps of this tile
          if (half == 0) {
            s += part[r] + net.b6;
#pragma unroll
            for (int c = 0; c < 28; ++c) s = fmaf(get_k(afeat, c), sw6[128 + c], s);
            s *= net.inv_scale;
            if (id >= 0) sdf_out[id] = negate ? -s : s;
          }
          asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");    // part[] / afeat reads done before restaging
        }
      }
    }
  } else if (warp < TC_EPI_WARPS + 4) {
    // ===================================== MMA issuers of tile pipeline g =====================================
    // A single thread sustains only one tcgen05.mma per ~100-120 clk (measured, tools/tc_bench.py) while the pipe
    // takes ~69 clk for M128 N128 K16: two issuer threads per tile.  Issuer A (sub 0) takes K chunks 0,2,4 and owns
    // the accumulate=false MMA; issuer B (sub 1) takes chunks 1,3 and starts once that first MMA has completed.
    const int g = (warp - TC_EPI_WARPS) >> 1;
    const int sub = (warp - TC_EPI_WARPS) & 1;
    const bool fast = (two_issuers & 2) != 0;   // single fp16 MMA per product (opt-in reduced-precision mode)
    two_issuers &= 1;
    if (tc::elect_one()) {
      const uint32_t idesc = tc::idesc_f16(128, 128, 0);
      const uint32_t ring = tc::smem_u32(smem + TS_RING);
      const uint32_t tD = tbase + g * 256, tAhi = tD + 128, tAlo = tD + 192;
      const uint32_t afeat = tc::smem_u32(smem + TS_AFEAT + g * TC_CHUNK_BYTES);
      const uint32_t ape = tc::smem_u32(smem + TS_APE + g * TC_CHUNK_BYTES);
      // constant parts of the shared-memory descriptors (K-major, no swizzle, LBO 2048, SBO 128)
      const uint64_t d0 = tc::smem_desc_kmajor(0, TC_LBO, TC_SBO);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32);
      const uint32_t ring_lo = (uint32_t)d0 | (ring >> 4);
      const uint32_t afeat_lo = (uint32_t)d0 | (afeat >> 4);
      const uint32_t ape_lo = (uint32_t)d0 | (ape >> 4);
      uint32_t ph_a = 0;
      int slot = 0;            // ring position / parity kept incrementally (a 64-bit % per chunk costs ~400 clk)
      uint32_t ring_par = 0;
      // The two tile pipelines have identical timing, so they would run in lockstep (both in their MMA phase,
      // then both in their epilogue phase).  Offset them once: Y's issuer starts after X has issued layer 1,
      // from then on X's epilogues overlap Y's MMAs and vice versa.
      if (g == 1) tc::mbar_wait(&bars->stagger, 0);
      uint32_t ph_f = 0;
      for (int64_t it = 0; it < my_pairs; ++it) {
        for (int l = 0; l < 6; ++l) {
          const int nch = (l == 0) ? 1 : 5;
          tc::mbar_wait(&bars->a_ready[g], ph_a & 1);
          ph_a++;
          tc::tc_fence_after();
          if (sub == 0) TRACE(100 * g + 30 + l);            // issuer: a_ready observed for layer l
          if (two_issuers && sub == 1 && l > 0) {
            tc::mbar_wait(&bars->f_done[g], ph_f & 1);
            ph_f++;
            tc::tc_fence_after();
          }
          for (int c = 0; c < nch; ++c, slot = (slot + 1 == TC_NSLOT) ? 0 : slot + 1, ring_par ^= (slot == 0)) {
            // one issuer (deterministic accumulation order) or chunks alternating between the two issuers
            if (two_issuers ? ((c & 1) != sub) : (sub != 0)) continue;
            tc::mbar_wait(&bars->w_full[slot], ring_par);
            // descriptor low words (address field is in 16-byte units): slot base, +256 per K step, +512 for lo
            const uint32_t w0 = ring_lo + slot * (TC_CHUNK_BYTES >> 4);
            if (l == 0 || c == 4) {
              const uint32_t a0 = (l == 0) ? ape_lo : afeat_lo;
              if (c == 0) {
                tc::mma_ss_w<false>(tD, a0, desc_hi, w0, desc_hi, idesc);
              } else {
                tc::mma_ss_w<true>(tD, a0, desc_hi, w0, desc_hi, idesc);
              }
              if (!fast) tc::mma_ss_w<true>(tD, a0 + 512, desc_hi, w0, desc_hi, idesc);
              if (!fast) tc::mma_ss_w<true>(tD, a0, desc_hi, w0 + 512, desc_hi, idesc);
              tc::mma_ss_w<true>(tD, a0 + 256, desc_hi, w0 + 256, desc_hi, idesc);
              if (!fast) tc::mma_ss_w<true>(tD, a0 + 768, desc_hi, w0 + 256, desc_hi, idesc);
              if (!fast) tc::mma_ss_w<true>(tD, a0 + 256, desc_hi, w0 + 768, desc_hi, idesc);
            } else {
              const uint32_t ah = tAhi + c * 16, al = tAlo + c * 16;
              if (c == 0) {
                tc::mma_ts_w<false>(tD, ah, w0, desc_hi, idesc);
                if (two_issuers) tc::mma_commit(&bars->f_done[g]);
              } else {
                tc::mma_ts_w<true>(tD, ah, w0, desc_hi, idesc);
              }
              if (!fast) tc::mma_ts_w<true>(tD, al, w0, desc_hi, idesc);
              if (!fast) tc::mma_ts_w<true>(tD, ah, w0 + 512, desc_hi, idesc);
              tc::mma_ts_w<true>(tD, ah + 8, w0 + 256, desc_hi, idesc);
              if (!fast) tc::mma_ts_w<true>(tD, al + 8, w0 + 256, desc_hi, idesc);
              if (!fast) tc::mma_ts_w<true>(tD, ah + 8, w0 + 768, desc_hi, idesc);
            }
            tc::mma_commit(&bars->w_empty[slot]);
          }
          tc::mma_commit(&bars->d_full[g]);
          if (g == 0 && sub == 0 && it == 0 && l == 1) tc::mbar_arrive(&bars->stagger);
          if (sub == 0) TRACE(100 * g + 40 + l);            // issuer: all MMAs of layer l issued
        }
      }
    }
  } else {
    // ===================================== weight loader (shared stream) =====================================
    if (lane == 0) {
      const int64_t total = my_pairs * TC_CHUNKS_FWD;
      uint8_t* ring = smem + TS_RING;
      int slot = 0, cid = 0;
      uint32_t par = 1;
      for (int64_t s = 0; s < total; ++s, slot = (slot + 1 == TC_NSLOT) ? 0 : slot + 1, par ^= (slot == 0),
                   cid = (cid + 1 == TC_CHUNKS_FWD) ? 0 : cid + 1) {
        if (s >= TC_NSLOT) tc::mbar_wait(&bars->w_empty[slot], par);
        tc::mbar_arrive_expect_tx(&bars->w_full[slot], TC_CHUNK_BYTES);
        tc::bulk_g2s(ring + slot * TC_CHUNK_BYTES, wblob + (size_t)cid * TC_CHUNK_BYTES,
                     TC_CHUNK_BYTES, &bars->w_full[slot]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS) tc::tmem_dealloc<512>(tbase);
}

// ---------------------------------------------------------------------------------------------
// host: tensor-core weight stream
// ---------------------------------------------------------------------------------------------
static inline uint16_t f2h_bits(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
static inline float h2f(uint16_t b) {
  __half h;
  memcpy(&h, &b, 2);
  return __half2float(h);
}

// W: folded fp32 weights per layer (out x in), skip scaling already applied
int surf_build_tc_weights(const std::vector<std::vector<float>>& W, const surf_net_inputs* in, surf_net* net,
                          cudaStream_t st, int (*dev_alloc)(surf_net*, void**, size_t)) {
  std::vector<uint16_t> blob((size_t)TC_CHUNKS_FWD * TC_CHUNK_BYTES / 2, 0);
  auto put = [&](int chunk, int n, int kk, float v) {    // kk in [0,32)
    const uint16_t hi = f2h_bits(v);
    const uint16_t lo = f2h_bits(v - h2f(hi));
    const size_t base = (size_t)chunk * (TC_CHUNK_BYTES / 2);
    const size_t off = (size_t)(kk >> 3) * 128 * 8 + (size_t)n * 8 + (kk & 7);
    blob[base + off] = hi;
    blob[base + TC_HALF_BYTES / 2 + off] = lo;
  };
  int chunk = 0;
  {   // lin0: K = 27 (+ bias at k = 27)
    const int O = in->out_dim[0], I = in->in_dim[0];
    for (int n = 0; n < O && n < 128; ++n) {
      for (int k = 0; k < I; ++k) put(chunk, n, k, W[0][(size_t)n * I + k]);
      put(chunk, n, 27, in->h_bias[0][n]);
    }
    chunk++;
  }
  for (int l = 1; l < 6; ++l) {
    const int O = in->out_dim[l], I = in->in_dim[l];
    for (int c = 0; c < 5; ++c) {
      for (int n = 0; n < O && n < 128; ++n) {
        for (int kk = 0; kk < 32; ++kk) {
          const int k = c * 32 + kk;
          if (k < I) put(chunk, n, kk, W[l][(size_t)n * I + k]);
          else if (k == 156) put(chunk, n, kk, in->h_bias[l][n]);
        }
      }
      chunk++;
    }
  }
  void* p = nullptr;
  int rc = dev_alloc(net, &p, blob.size() * 2);
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice, st));
  SURF_CUDA(cudaStreamSynchronize(st));
  net->tc_blob = (const uint8_t*)p;
  return 0;
}

int launch_sdf_tc_fwd(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_sdf, bool negate,
                      cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SURF_CUDA(cudaFuncSetAttribute(k_sdf_tc_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_TOTAL));
    attr_set = true;
  }
  if (src.n <= 0) return 0;
  const int64_t pairs = ((src.n + 127) / 128 + 1) / 2;
  const int grid = (int)(pairs < n->n_sm ? pairs : n->n_sm);
  surf_time_begin(1, st);
  k_sdf_tc_fwd<<<grid, TC_THREADS, TS_TOTAL, st>>>(s->dev, n->dev, src, n->tc_blob, d_sdf, negate ? 1 : 0,
                                                   (surf_mlp_mode() == 2 ? 1 : 0) | (surf_mlp_mode() == 4 ? 2 : 0));
  surf_time_end(1, st);
  SURF_LAUNCH_CHECK();
  return 0;
}

#ifdef TC_TRACE
extern "C" int surf_tc_trace_read(long long* h_out, int max_events) {
  int n = 0;
  cudaMemcpyFromSymbol(&n, g_tc_trace_n, sizeof(int));
  if (n > max_events) n = max_events;
  if (n > 2048) n = 2048;
  cudaMemcpyFromSymbol(h_out, g_tc_trace, sizeof(long long) * 2 * n);
  int zero = 0;
  cudaMemcpyToSymbol(g_tc_trace_n, &zero, sizeof(int));
  return n;
}
#endif
