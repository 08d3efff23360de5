// EXPERIMENT (round 2) — NOT part of the product library; kept with its measurements (tools/experiments/README.md).
// Builds against the product headers when dropped into surf_b200/csrc/ in place of sdf_tc2.cu (launch_sdf_tc2 ->
// launch_sdf_tc, surf_net needs an `int32_t* tc_rows` member, tc_common.cuh an un-hinted mbar_wait_fast).
//
// K2a + K3, tensor-core edition with TWO 128-point tiles in flight per CTA ("ping-pong"): SDF MLP forward + analytic
// input gradient.
//   reference: SDFNetworkSparse.sdf / .gradient (sdf_network.py:95-141), lookup_sparse_volume (projector.py:217-390)
//
// The layer chain of one tile is strictly serial — MMAs of layer p, activation epilogue of layer p, MMAs of layer p+1 —
// so with one tile per CTA either the tensor pipe or the 16 epilogue warps idle (round 1: tensor pipe 38 % active, 16 %
// of the warp samples waiting for the MMA-done barrier).  Here a CTA works on two tiles A / B that are half a step
// apart:
//        tensor pipe :  MMA(A,p)   MMA(B,p)   MMA(A,p+1)  MMA(B,p+1) ...
//        epilogue    :  EPI(B,p-1) EPI(A,p)   EPI(B,p)    EPI(A,p+1) ...
// and both units stay busy.  What makes two tiles fit into the 512 TMEM columns:
//   * the epilogue works IN PLACE: a warp reads its 16 accumulator columns (fp32) and writes the fp16 hi (8 columns) |
//     lo (8 columns) A operand of the next layer over them, so an accumulator and the operand made from it share 128
//     columns.  Three 128-column regions rotate: step k = 2 p + x (x = tile parity) accumulates into region k % 3, reads
//     its A operand from region (k - 2) % 3 while the other tile's epilogue runs in region (k - 1) % 3;
//   * the reverse pass no longer widens every accumulator to N = 160: the feature-gradient columns accumulate ACROSS
//     the reverse layers in one persistent 32-column accumulator per tile (DFEAT), the positional-encoding gradient
//     of the skip layer and of lin0 in another (DPE); the helper warps read both straight from TMEM.
// One thread issues every MMA in a fixed order into fixed accumulators: results are bitwise deterministic.
// Weights: fp16 hi | lo chunks streamed L2 -> smem ring by bulk copies; every fp32 product is hi*hi + lo*hi + hi*lo
// (3 MMAs, fp32-grade) or hi*hi only (opt-in 1e-2 mode).  softplus' for the reverse pass: 16-bit codes in a per-CTA
// L2-resident scratch.
#include <math.h>
#include <string.h>

#include <vector>

#include "surf_internal.cuh"
#include "tc_common.cuh"

// The same fetch split in two steps for latency-bound callers (the helper warps of sdf_tc3.cu: one point per thread,
// nothing else to hide a load behind).  sparse_rows issues the 8 index loads of a level; sparse_accum loads the voxel
// rows BRANCH-FREE in two batches of four corners (8 x 128-bit loads in flight) — a missing corner (row < 0) reads
// row 0 with weight 0, which adds +-0 and leaves every result bit-identical to sparse_level (`continue` on a missing
// corner serialises the eight row fetches into eight dependent L2 round trips per level).
__device__ __forceinline__ void sparse_rows(const DevScene& sc, int l, float px, float py, float pz, int32_t (&rows)[8]) {
  const int N = sc.dim[l];
  const float vs = sc.voxel[l];
  const float fx0 = floorf(__fdiv_rn(__fadd_rn(pz, 1.0f), vs));
  const float fy0 = floorf(__fdiv_rn(__fadd_rn(py, 1.0f), vs));
  const float fz0 = floorf(__fdiv_rn(__fadd_rn(px, 1.0f), vs));
  const float hi = (float)(N - 1);
  const int x0 = (int)fminf(fmaxf(fx0, 0.f), hi), x1 = (int)fminf(fmaxf(fx0 + 1.0f, 0.f), hi);
  const int y0 = (int)fminf(fmaxf(fy0, 0.f), hi), y1 = (int)fminf(fmaxf(fy0 + 1.0f, 0.f), hi);
  const int z0 = (int)fminf(fmaxf(fz0, 0.f), hi), z1 = (int)fminf(fmaxf(fz0 + 1.0f, 0.f), hi);
  const int32_t* __restrict__ idx = sc.index[l];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int xi = (c & 1) ? x1 : x0, yi = (c & 2) ? y1 : y0, zi = (c & 4) ? z1 : z0;
    rows[c] = __ldg(idx + ((size_t)zi * N + yi) * N + xi);
  }
}

template <int MODE>
__device__ __forceinline__ void sparse_accum(const DevScene& sc, int l, float px, float py, float pz,
                                             const int32_t (&rows)[8], const float* g, float* out) {
  const float vs = sc.voxel[l];
  const float cx = __fdiv_rn(__fadd_rn(pz, 1.0f), vs);
  const float cy = __fdiv_rn(__fadd_rn(py, 1.0f), vs);
  const float cz = __fdiv_rn(__fadd_rn(px, 1.0f), vs);
  const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
  const float wx1 = __fsub_rn(cx, fx0), wx0 = __fsub_rn(fx0 + 1.0f, cx);
  const float wy1 = __fsub_rn(cy, fy0), wy0 = __fsub_rn(fy0 + 1.0f, cy);
  const float wz1 = __fsub_rn(cz, fz0), wz0 = __fsub_rn(fz0 + 1.0f, cz);
  const float4* __restrict__ vol = sc.vol8[l];
  if (MODE == 0) {
#pragma unroll
    for (int c = 0; c < 7; ++c) out[c] = 0.f;
  } else {
    out[0] = out[1] = out[2] = 0.f;
  }
  float gx = 0.f, gy = 0.f, gz = 0.f;
  bool any = false;
#pragma unroll
  for (int c = 0; c < 8; ++c) any = any || rows[c] >= 0;
  if (any) {                       // (a level without a single voxel has no volume buffer to read row 0 from)
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      float4 a[4], b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int32_t rw = rows[hb * 4 + j] < 0 ? 0 : rows[hb * 4 + j];
        a[j] = __ldg(vol + (size_t)rw * 2);
        b[j] = __ldg(vol + (size_t)rw * 2 + 1);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // order bnw,bne,bsw,bse,fnw,fne,fsw,fse (projector.py:360-371)
        const int c = hb * 4 + j;
        const float m = rows[c] < 0 ? 0.f : 1.f;
        const float wx = (c & 1) ? wx1 : wx0, wy = (c & 2) ? wy1 : wy0, wz = (c & 4) ? wz1 : wz0;
        if (MODE == 0) {
          const float w = __fmul_rn(__fmul_rn(wx, wy), wz) * m;
          out[0] += a[j].x * w; out[1] += a[j].y * w; out[2] += a[j].z * w; out[3] += a[j].w * w;
          out[4] += b[j].x * w; out[5] += b[j].y * w; out[6] += b[j].z * w;
        } else {
          const float s = (a[j].x * g[0] + a[j].y * g[1] + a[j].z * g[2] + a[j].w * g[3] + b[j].x * g[4] + b[j].y * g[5] +
                           b[j].z * g[6]) * m;
          gx += s * (((c & 1) ? 1.f : -1.f) * wy * wz);
          gy += s * (((c & 2) ? 1.f : -1.f) * wx * wz);
          gz += s * (((c & 4) ? 1.f : -1.f) * wx * wy);
        }
      }
    }
  }
  if (MODE == 1) {
    const float inv = 1.0f / vs;
    out[0] = gz * inv;   // d/d world x  (grid z)
    out[1] = gy * inv;
    out[2] = gx * inv;   // d/d world z  (grid x)
  }
}



#define T3_EPI_WARPS 16
#define T3_EPI_THREADS (T3_EPI_WARPS * 32)
#define T3_HELP_WARPS 4
#define T3_THREADS ((T3_EPI_WARPS + 2 + T3_HELP_WARPS) * 32)     // + 1 MMA issuer + 1 weight loader + helpers
#define T3_SLOT_BYTES 20480                      // 160 rows x K = 32 x (hi + lo)
#define T3_NSLOT 4
#define T3_NSTAGE 4                              // staged tiles (2 in flight + 2 being prepared)
#define T3_SCRATCH_U4 (5 * 4 * T3_EPI_THREADS + 5 * T3_EPI_THREADS / 4)   // per (CTA, tile parity) scratch, in uint4

// TMEM columns
#define T3_REGION(i) ((uint32_t)(i) * 128u)
#define T3_DFEAT(x) (384u + 32u * (uint32_t)(x))
#define T3_DPE(x) (448u + 32u * (uint32_t)(x))
#define T3_PE_SHIFT 5                            // PE index k lives in DPE column k + 5 (rows 96.. of the skip layer)

// dynamic smem (bytes)
#define S3_RING 0
#define S3_STAGE (S3_RING + T3_NSLOT * T3_SLOT_BYTES)      // T3_NSTAGE x [features hi 8 KB | lo 8 KB | PE hi 8 KB | lo 8 KB]
#define S3_STAGE_BYTES 32768
#define S3_W6 (S3_STAGE + T3_NSTAGE * S3_STAGE_BYTES)      // 160 floats
#define S3_PART (S3_W6 + 640)                              // [4][128] floats
#define S3_BAR (S3_PART + 2048)
#define S3_TOTAL (S3_BAR + 256)

// timing experiments (tools/sdf_bench.py; only in a -DT3_DEBUG build, results are then invalid): flag bits
//   4 no tcgen05.mma, 8 no activation math, 16 no softplus' scratch traffic, 32 no weight traffic, 64 no gathers,
//   128 no DFEAT / DPE MMAs, 256 no final gradient, 512 suspend-hinted waits on the critical path
#ifdef T3_DEBUG
#define T3_DBG(bit) ((flags & (bit)) != 0)
#else
#define T3_DBG(bit) false
#endif

#ifdef T3_DEBUG
// clock accounting of CTA 0's roles (tools/sdf_bench.py): [0] issuer total, [1] issuer wait a_ready, [2] issuer wait
// weights, [3] issuer wait stage_ready, [4] issuer wait finish_done, [5] epilogue warp 0 total, [6] its wait d_full,
// [7] helper warp total, [8] helper stage, [9] helper wait tile_done, [10] helper finish, [11] loader wait w_empty
__device__ unsigned long long g_t3_prof[16];
#define T3_PROF_DECL unsigned long long prof_acc[4] = {0ull, 0ull, 0ull, 0ull}; const long long prof_t0 = clock64();
#define T3_PROF_BEGIN const long long prof_b = clock64();
#define T3_PROF_END(i) prof_acc[i] += (unsigned long long)(clock64() - prof_b);
#define T3_PROF_FLUSH(cond, base, n)                                                        \
  if (blockIdx.x == 0 && (cond)) {                                                          \
    atomicAdd(&g_t3_prof[base], (unsigned long long)(clock64() - prof_t0));                  \
    for (int _i = 0; _i < (n); ++_i) atomicAdd(&g_t3_prof[(base) + 1 + _i], prof_acc[_i]);  \
  }
#else
#define T3_PROF_DECL
#define T3_PROF_BEGIN
#define T3_PROF_END(i)
#define T3_PROF_FLUSH(cond, base, n)
#endif

struct T3Bars {
  uint64_t w_full[T3_NSLOT];
  uint64_t w_empty[T3_NSLOT];
  uint64_t d_full[2];             // all MMAs of a layer phase of the tile of parity x are done (tcgen05.commit)
  uint64_t a_ready[2];            // the epilogue of a phase is done with its region (next A operand written)
  uint64_t stage_ready[T3_NSTAGE];  // smem operands of a tile staged: one arrival per helper warp
  uint64_t tile_done[2];          // GRAD: phase 11 committed; else: the head epilogue is done (one arrival per warp)
  uint64_t finish_done[2];        // the helpers have read DFEAT / DPE of the tile of parity x
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
// softplus(beta = 100), branch-free: max(z, 0) + log1p(e) / 100 with e = exp(-|100 z|) in (0, 1].
// softplus'(z) = sigmoid(100 z) = z >= 0 ? 1 / (1 + e) : 1 - 1 / (1 + e).  For the reverse pass the forward epilogue
// parks e as a 16-bit fixed-point code plus the sign of z (no F2I: min(e, 1 - 2^-16) + 128.0f has ulp 2^-16, the
// adder's round-to-nearest leaves round(e * 65536) in the low 16 bits of the word).
__device__ __forceinline__ float t3_softplus(float z, float& e) {
  e = t3_ex2(fabsf(z) * -144.26950408889634f);
  return fmaf(t3_lg2(1.0f + e), 0.0069314718055994531f, fmaxf(z, 0.f));
}
__device__ __forceinline__ uint32_t t3_code_word(float e) { return __float_as_uint(fminf(e, 0.9999847412109375f) + 128.0f); }
template <int T>
__device__ __forceinline__ float t3_decode_u(uint32_t pair) {     // pair word (two codes) -> u = 1 + e of element T
  return __uint_as_float(__byte_perm(pair, 0x43000000u, T ? 0x7632 : 0x7610)) - 127.0f;
}

__device__ __forceinline__ float t3_get_k(const uint8_t* base, int r, int k) {   // hi + lo of a staged smem operand
  const uint32_t off = (uint32_t)(k >> 3) * 2048u + r * 16 + (k & 7) * 2;
  return __half2float(*reinterpret_cast<const __half*>(base + off)) +
         __half2float(*reinterpret_cast<const __half*>(base + 8192 + off));
}

// wait for the outstanding tcgen05.ld; the registers it fills are operands so that no use can be scheduled above it
__device__ __forceinline__ void t3_wait_ld(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

template <bool ACC>
__device__ __forceinline__ void t3_ts(bool skip, uint32_t d, uint32_t a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  if (!skip) tc::mma_ts_w<ACC>(d, a, b_lo, b_hi, idesc);
}
template <bool ACC>
__device__ __forceinline__ void t3_ss(bool skip, uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                      uint32_t idesc) {
  if (!skip) tc::mma_ss_w<ACC>(d, a_lo, a_hi, b_lo, b_hi, idesc);
}

struct T3Epi {
  int flags;
  uint32_t tl;            // TMEM base of my lane quarter
  int part, r, te;
  const uint8_t* ape;     // staged positional encoding of the current tile
  const float* sw6;
  float inv_scale;
  uint4* scratch;         // of the current tile parity
  uint32_t* sgn_scratch;
};

// 16 values -> fp16 hi (8 words) | lo (8 words), written over the 16 accumulator columns they came from
__device__ __forceinline__ void t3_store_a(uint32_t col, const float (&h)[16]) {
  uint32_t w[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) tc::split2(h[2 * j], h[2 * j + 1], w[j], w[8 + j]);
  tc::tmem_st16(col, w);
}

// Forward epilogue of one 16-column block (columns cb .. cb+15 of the layer): h = softplus(z) -> next A operand.
// PE_FROM: first of my 16 columns that is a positional-encoding input of the skip layer (16 = none).
// HEAD: lin5 -> SDF head partial sum and (GRAD) h = delta5 = w6 / scale * softplus', the first reverse A operand.
template <bool GRAD, bool HEAD>
__device__ __forceinline__ void t3_fwd_block(const T3Epi& c, int cb, int pe_from, const uint32_t (&d)[16], float (&h)[16],
                                             uint4 (&spw)[2], uint32_t& sgn, float& head) {
  const int flags = c.flags;
  (void)flags;
  uint32_t cw[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    cw[n] = 0u;
    const float z = __uint_as_float(d[n]);
    if (T3_DBG(8)) {
      h[n] = z;
    } else if (!HEAD && n >= pe_from) {
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
        sgn = __funnelshift_l(__float_as_uint(z), sgn, 1);      // element e of my 32 ends at bit 31 - e
      }
    }
  }
  if (GRAD && !HEAD) {
    spw[0] = make_uint4(__byte_perm(cw[0], cw[1], 0x5410), __byte_perm(cw[2], cw[3], 0x5410),
                        __byte_perm(cw[4], cw[5], 0x5410), __byte_perm(cw[6], cw[7], 0x5410));
    spw[1] = make_uint4(__byte_perm(cw[8], cw[9], 0x5410), __byte_perm(cw[10], cw[11], 0x5410),
                        __byte_perm(cw[12], cw[13], 0x5410), __byte_perm(cw[14], cw[15], 0x5410));
  }
}

// Forward phase l of one tile: accumulator region `reg` -> (in place) the A operand of layer l + 1.
template <bool GRAD, bool HEAD>
__device__ __forceinline__ void t3_fwd_layer(const T3Epi& c, uint32_t reg, int l, bool skip_next, float& head) {
  const uint32_t col0 = c.tl + reg + c.part * 32;
  const int flags = c.flags;
  (void)flags;
  uint32_t sgn = 0;
  uint32_t da[16], db[16];
  tc::tmem_ld16(col0, da);
  tc::tmem_ld16(col0 + 16, db);
  constexpr bool STORE = !HEAD || GRAD;
  {
    const int cb = c.part * 32;
    t3_wait_ld(da);              // (both loads have landed: tcgen05.wait::ld waits for all of them)
    float h[16];
    uint4 spw[2];
    const int pe_from = (skip_next && c.part == 3) ? 5 : 16;
    t3_fwd_block<GRAD, HEAD>(c, cb, pe_from, da, h, spw, sgn, head);
    if (STORE) t3_store_a(col0, h);
    if (GRAD && !HEAD && !T3_DBG(16)) {
      c.scratch[(size_t)(l * 4 + 0) * T3_EPI_THREADS + c.te] = spw[0];
      c.scratch[(size_t)(l * 4 + 1) * T3_EPI_THREADS + c.te] = spw[1];
    }
  }
  {
    const int cb = c.part * 32 + 16;
    t3_wait_ld(db);              // (already landed; pins the uses of db below the wait)
    float h[16];
    uint4 spw[2];
    const int pe_from = (skip_next && c.part == 3) ? 0 : 16;
    t3_fwd_block<GRAD, HEAD>(c, cb, pe_from, db, h, spw, sgn, head);
    if (STORE) t3_store_a(col0 + 16, h);
    if (GRAD && !HEAD && !T3_DBG(16)) {
      c.scratch[(size_t)(l * 4 + 2) * T3_EPI_THREADS + c.te] = spw[0];
      c.scratch[(size_t)(l * 4 + 3) * T3_EPI_THREADS + c.te] = spw[1];
    }
  }
  if (GRAD && !HEAD) c.sgn_scratch[(size_t)l * T3_EPI_THREADS + c.te] = sgn;
}

// Reverse phase: region `reg` holds d sdf / d (input of lin_l) (128 hidden columns); v = delta_{l-1} = D * softplus'
// (z_{l-1}) -> in place, the A operand of the next reverse layer.  skip_pe: columns 101.. are the positional-encoding
// inputs of the skip layer — their gradient was accumulated into DPE by the tensor core, v = 0 here.
__device__ __forceinline__ void t3_bwd_layer(const T3Epi& c, uint32_t reg, bool skip_pe, const uint4 (&codes)[4],
                                             uint32_t sgn) {
  const uint32_t col0 = c.tl + reg + c.part * 32;
  const int flags = c.flags;
  (void)flags;
  uint32_t da[16], db[16];
  tc::tmem_ld16(col0, da);
  tc::tmem_ld16(col0 + 16, db);
  t3_wait_ld(da);
  t3_wait_ld(db);
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const uint32_t (&d)[16] = b ? db : da;
    const uint32_t sp[8] = {codes[2 * b].x, codes[2 * b].y, codes[2 * b].z, codes[2 * b].w,
                            codes[2 * b + 1].x, codes[2 * b + 1].y, codes[2 * b + 1].z, codes[2 * b + 1].w};
    const int pe_from = (skip_pe && c.part == 3) ? (b ? 0 : 5) : 16;
    float v[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const float g = __uint_as_float(d[n]);
      if (T3_DBG(8)) {
        v[n] = g;
      } else if (n >= pe_from) {
        v[n] = 0.f;
      } else {
        const float rr = t3_rcp((n & 1) ? t3_decode_u<1>(sp[n >> 1]) : t3_decode_u<0>(sp[n >> 1]));
        const bool neg = ((sgn >> (31 - (b * 16 + n))) & 1u) != 0u;
        v[n] = g * (neg ? 1.0f - rr : rr);
      }
    }
    t3_store_a(col0 + b * 16, v);
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(T3_THREADS, 1)
k_sdf_pp(const DevScene sc, const DevNet net, const PointSource src, const uint8_t* __restrict__ wblob,
         const T1Stream stream, float* __restrict__ sdf_out, float* __restrict__ grad_out,
         uint4* __restrict__ scratch_all, int32_t* __restrict__ rows_all, int flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  T3Bars* bars = reinterpret_cast<T3Bars*>(smem + S3_BAR);
  float* sw6 = reinterpret_cast<float*>(smem + S3_W6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NPH = GRAD ? 12 : 6;           // MMA phases per tile
  constexpr int NEPI = GRAD ? 11 : 6;          // epilogue phases per tile

  int64_t n_total = src.n;
  if (src.count) {
    const int64_t c = *src.count;
    n_total = c < n_total ? c : n_total;
  }
  const int64_t n_tiles = (n_total + 127) / 128;
  int64_t my_tiles = 0;
  if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t n_pairs = (my_tiles + 1) >> 1;

  if (warp == T3_EPI_WARPS) tc::tmem_alloc<512>(&bars->tmem_base);
  if (tid == 0) {
    for (int i = 0; i < T3_NSLOT; ++i) {
      tc::mbar_init(&bars->w_full[i], 1);
      tc::mbar_init(&bars->w_empty[i], 1);
    }
    for (int x = 0; x < 2; ++x) {
      tc::mbar_init(&bars->d_full[x], 1);
      tc::mbar_init(&bars->a_ready[x], T3_EPI_WARPS);
      tc::mbar_init(&bars->tile_done[x], GRAD ? 1 : T3_EPI_WARPS);
      tc::mbar_init(&bars->finish_done[x], T3_HELP_WARPS);
    }
    for (int b = 0; b < T3_NSTAGE; ++b) tc::mbar_init(&bars->stage_ready[b], T3_HELP_WARPS);
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

  if (warp < T3_EPI_WARPS) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, part = warp >> 2;
    const int r = q * 32 + lane;                 // row = point = TMEM lane
    float* s_part = reinterpret_cast<float*>(smem + S3_PART);
    const int te = warp * 32 + lane;             // 0..511
    uint32_t ph_d[2] = {0u, 0u};
    auto epi_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory"); };

    T3_PROF_DECL
    T3Epi ec;
    ec.flags = flags;
    ec.tl = tbase + ((uint32_t)(q * 32) << 16); ec.part = part; ec.r = r; ec.te = te; ec.ape = nullptr; ec.sw6 = sw6;
    ec.inv_scale = net.inv_scale; ec.scratch = nullptr; ec.sgn_scratch = nullptr;
    uint4* scratch_cta = scratch_all + (size_t)blockIdx.x * 2 * T3_SCRATCH_U4;

    for (int64_t j = 0; j < n_pairs; ++j) {
#pragma unroll 1
      for (int p = 0; p < NEPI; ++p) {
#pragma unroll 1
        for (int x = 0; x < 2; ++x) {
          const int64_t it = 2 * j + x;
          if (it >= my_tiles) continue;
          const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
          const int buf = (int)(it & (T3_NSTAGE - 1));
          const uint8_t* afeat = smem + S3_STAGE + buf * S3_STAGE_BYTES;
          ec.ape = afeat + 16384;
          ec.scratch = scratch_cta + (size_t)x * T3_SCRATCH_U4;
          ec.sgn_scratch = reinterpret_cast<uint32_t*>(ec.scratch + 5 * 4 * T3_EPI_THREADS);
          const uint32_t reg = T3_REGION((2 * p + x) % 3);
          // reverse phases: the softplus' codes come from L2 — request them before waiting for the accumulator
          uint4 codes[4] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u),
                            make_uint4(0u, 0u, 0u, 0u)};
          uint32_t sgn = 0;
          if (GRAD && p >= 6 && !T3_DBG(16)) {
            const int lsrc = 10 - p;                          // reverse of lin_l, l = 11 - p; codes of layer l - 1
            const uint4* cp = ec.scratch + (size_t)(lsrc * 4) * T3_EPI_THREADS + te;
#pragma unroll
            for (int i = 0; i < 4; ++i) codes[i] = cp[(size_t)i * T3_EPI_THREADS];
            sgn = ec.sgn_scratch[(size_t)lsrc * T3_EPI_THREADS + te];
          }
          if (p == 0) tc::mbar_wait(&bars->stage_ready[buf], (uint32_t)((it >> 2) & 1));   // acquire the helpers' smem writes
          {
            T3_PROF_BEGIN
            if (T3_DBG(512)) tc::mbar_wait(&bars->d_full[x], ph_d[x] & 1); else tc::mbar_wait_fast(&bars->d_full[x], ph_d[x] & 1);
            T3_PROF_END(0)
          }
          ph_d[x]++;
          tc::tc_fence_after();
          if (p < 5) {
            float head = 0.f;
            t3_fwd_layer<GRAD, false>(ec, reg, p, p + 1 == net.skip_layer, head);
          } else if (p == 5) {
            float head = 0.f;
            t3_fwd_layer<GRAD, true>(ec, reg, p, false, head);
            s_part[part * 128 + r] = head;
            epi_bar();
            if (part == 0) {
              float s = s_part[r] + s_part[128 + r] + s_part[256 + r] + s_part[384 + r] + net.b6;
#pragma unroll
              for (int c = 0; c < 28; ++c) s = fmaf(t3_get_k(afeat, r, c), sw6[128 + c], s);
              s *= net.inv_scale;
              const int64_t i = tile * 128 + r;
              if (i < n_total) {
                const int64_t id = src.list ? (int64_t)src.list[i] : i;
                sdf_out[id] = (flags & 1) ? -s : s;
              }
            }
            epi_bar();            // s_part is reused by the other tile's head one step later
          } else {
            t3_bwd_layer(ec, reg, (11 - p) == net.skip_layer, codes, sgn);
          }
          // my part of the region is final: A operand stores landed (or, head without gradient: accumulator read)
          tc::tmem_wait_st();
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            tc::mbar_arrive(&bars->a_ready[x]);
            if (!GRAD && p == 5) tc::mbar_arrive(&bars->tile_done[x]);
          }
        }
      }
    }
    T3_PROF_FLUSH(warp == 0 && lane == 0, 5, 1)
  } else if (warp == T3_EPI_WARPS) {
    // =============================== the MMA issuer ===============================
    const bool fast = (flags & 2) != 0;       // single fp16 MMA per product (opt-in reduced-precision mode)
    const bool nomma = T3_DBG(4), nosmall = T3_DBG(128);
    if (tc::elect_one()) {
      T3_PROF_DECL
      const uint32_t ring = tc::smem_u32(smem + S3_RING);
      const uint32_t id128 = tc::idesc_f16(128, 128, 0), id32 = tc::idesc_f16(128, 32, 0);
      // descriptor constant parts: SBO 128; LBO = rows * 16
      const uint64_t d128 = tc::smem_desc_kmajor(0, 2048, 128), d160 = tc::smem_desc_kmajor(0, 2560, 128),
                     d32 = tc::smem_desc_kmajor(0, 512, 128);
      const uint32_t stage_lo0 = (uint32_t)d128 | (tc::smem_u32(smem + S3_STAGE) >> 4);
      const uint32_t dh128 = (uint32_t)(d128 >> 32), dh160 = (uint32_t)(d160 >> 32), dh32 = (uint32_t)(d32 >> 32);
      uint32_t ph_a[2] = {0u, 0u};
      int slot = 0;
      uint32_t ring_par = 0;
      auto next_chunk = [&]() -> uint32_t {      // wait for the next weight chunk; its smem address in 16-byte units
        T3_PROF_BEGIN
        if (T3_DBG(512)) tc::mbar_wait(&bars->w_full[slot], ring_par); else tc::mbar_wait_fast(&bars->w_full[slot], ring_par);
        T3_PROF_END(1)
        return (ring + slot * T3_SLOT_BYTES) >> 4;
      };
      auto release_chunk = [&]() {
        tc::mma_commit(&bars->w_empty[slot]);
        slot = (slot + 1 == T3_NSLOT) ? 0 : slot + 1;
        ring_par ^= (slot == 0);
      };
      auto wait_a = [&](int x) {
        T3_PROF_BEGIN
        if (T3_DBG(512)) tc::mbar_wait(&bars->a_ready[x], ph_a[x] & 1); else tc::mbar_wait_fast(&bars->a_ready[x], ph_a[x] & 1);
        T3_PROF_END(0)
        ph_a[x]++;
        tc::tc_fence_after();
      };
      for (int64_t j = 0; j < n_pairs; ++j) {
#pragma unroll 1
        for (int p = 0; p < NPH; ++p) {
#pragma unroll 1
          for (int x = 0; x < 2; ++x) {
            const int64_t it = 2 * j + x;
            if (it >= my_tiles) continue;
            const int buf = (int)(it & (T3_NSTAGE - 1));
            const uint32_t afeat_lo = stage_lo0 + buf * (S3_STAGE_BYTES >> 4), ape_lo = afeat_lo + 1024;
            const uint32_t tD = tbase + T3_REGION((2 * p + x) % 3);
            const uint32_t tA = tbase + T3_REGION((2 * p + x + 1) % 3);          // = region of step 2 (p - 1) + x
            if (p == 0) {
              // ---- lin0 on the positional encoding (A in smem) ----
              {
                T3_PROF_BEGIN
                tc::mbar_wait(&bars->stage_ready[buf], (uint32_t)((it >> 2) & 1));
                T3_PROF_END(2)
              }
              // without the reverse pass the previous tile of this parity ends with an epilogue that only READS its
              // region: it must be through before anything rotates back onto that region
              if (!GRAD && it >= 2) wait_a(x);
              tc::tc_fence_after();
              const uint32_t w0 = (uint32_t)d128 | next_chunk(), dh = dh128, a0 = ape_lo;
              // the small correction products (lo x hi, hi x lo) go first: the tensor core's fp32 accumulation
              // truncates, so they are added while the accumulator is still small
              if (!fast) {
                t3_ss<false>(nomma, tD, a0 + 512, dh, w0, dh, id128);
                t3_ss<true>(nomma, tD, a0, dh, w0 + 512, dh, id128);
                t3_ss<true>(nomma, tD, a0 + 768, dh, w0 + 256, dh, id128);
                t3_ss<true>(nomma, tD, a0 + 256, dh, w0 + 768, dh, id128);
                t3_ss<true>(nomma, tD, a0, dh, w0, dh, id128);
              } else {
                t3_ss<false>(nomma, tD, a0, dh, w0, dh, id128);
              }
              t3_ss<true>(nomma, tD, a0 + 256, dh, w0 + 256, dh, id128);
              release_chunk();
              tc::mma_commit(&bars->d_full[x]);
            } else if (p < 6) {
              // ---- lin1..lin5: feature | bias columns (A in smem, independent of the previous layer) first ----
              const uint32_t dh = dh128;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint32_t w0 = (uint32_t)d128 | next_chunk();
                const uint32_t a0 = afeat_lo + h * 256;
                if (!fast) {
                  if (h == 0) t3_ss<false>(nomma, tD, a0 + 512, dh, w0, dh, id128); else t3_ss<true>(nomma, tD, a0 + 512, dh, w0, dh, id128);
                  t3_ss<true>(nomma, tD, a0, dh, w0 + 256, dh, id128);
                  t3_ss<true>(nomma, tD, a0, dh, w0, dh, id128);
                } else {
                  if (h == 0) t3_ss<false>(nomma, tD, a0, dh, w0, dh, id128); else t3_ss<true>(nomma, tD, a0, dh, w0, dh, id128);
                }
                release_chunk();
              }
              wait_a(x);
#pragma unroll 1
              for (int c = 0; c < 4; ++c) {        // hidden columns: K chunk c = A blocks 2c, 2c+1 (hi at +0, lo at +8)
                const uint32_t w0 = (uint32_t)d128 | next_chunk();
                const uint32_t ah = tA + c * 32, al = ah + 8;
                if (!fast) {
                  t3_ts<true>(nomma, tD, al, w0, dh, id128);
                  t3_ts<true>(nomma, tD, ah, w0 + 512, dh, id128);
                  t3_ts<true>(nomma, tD, al + 16, w0 + 256, dh, id128);
                  t3_ts<true>(nomma, tD, ah + 16, w0 + 768, dh, id128);
                }
                t3_ts<true>(nomma, tD, ah, w0, dh, id128);
                t3_ts<true>(nomma, tD, ah + 16, w0 + 256, dh, id128);
                release_chunk();
              }
              tc::mma_commit(&bars->d_full[x]);
            } else if (p < 11) {
              // ---- reverse of lin5..lin1: B[n = input index (160 rows)][k = output index], four K = 32 chunks.
              //      rows 0..127 -> hidden gradient (region), rows 128..159 -> DFEAT (accumulates over the layers),
              //      skip layer: rows 96..127 -> DPE ----
              wait_a(x);
              if (p == 6 && it >= 2) {
                T3_PROF_BEGIN
                tc::mbar_wait(&bars->finish_done[x], (uint32_t)(((it >> 1) - 1) & 1));
                T3_PROF_END(3)
                tc::tc_fence_after();
              }
              const uint32_t dh = dh160;
              const uint32_t tF = tbase + T3_DFEAT(x), tP = tbase + T3_DPE(x);
              const bool pe = (11 - p) == net.skip_layer;
#pragma unroll 1
              for (int c = 0; c < 4; ++c) {
                const uint32_t w0 = (uint32_t)d160 | next_chunk();
                const uint32_t ah = tA + c * 32, al = ah + 8;
                if (!fast) {
                  if (c == 0) t3_ts<false>(nomma, tD, al, w0, dh, id128); else t3_ts<true>(nomma, tD, al, w0, dh, id128);
                  t3_ts<true>(nomma, tD, ah, w0 + 640, dh, id128);
                  t3_ts<true>(nomma, tD, al + 16, w0 + 320, dh, id128);
                  t3_ts<true>(nomma, tD, ah + 16, w0 + 960, dh, id128);
                  t3_ts<true>(nomma, tD, ah, w0, dh, id128);
                } else {
                  if (c == 0) t3_ts<false>(nomma, tD, ah, w0, dh, id128); else t3_ts<true>(nomma, tD, ah, w0, dh, id128);
                }
                t3_ts<true>(nomma, tD, ah + 16, w0 + 320, dh, id128);
                {   // feature-gradient columns: rows 128..159 of the same chunk
                  const uint32_t wf = w0 + 128;
                  if (!fast) {
                    if (p == 6 && c == 0) t3_ts<false>(nomma || nosmall, tF, al, wf, dh, id32); else t3_ts<true>(nomma || nosmall, tF, al, wf, dh, id32);
                    t3_ts<true>(nomma || nosmall, tF, ah, wf + 640, dh, id32);
                    t3_ts<true>(nomma || nosmall, tF, al + 16, wf + 320, dh, id32);
                    t3_ts<true>(nomma || nosmall, tF, ah + 16, wf + 960, dh, id32);
                    t3_ts<true>(nomma || nosmall, tF, ah, wf, dh, id32);
                  } else {
                    if (p == 6 && c == 0) t3_ts<false>(nomma || nosmall, tF, ah, wf, dh, id32); else t3_ts<true>(nomma || nosmall, tF, ah, wf, dh, id32);
                  }
                  t3_ts<true>(nomma || nosmall, tF, ah + 16, wf + 320, dh, id32);
                }
                if (pe) {   // positional-encoding columns of the skip layer: rows 96..127
                  const uint32_t wp = w0 + 96;
                  if (!fast) {
                    if (c == 0) t3_ts<false>(nomma || nosmall, tP, al, wp, dh, id32); else t3_ts<true>(nomma || nosmall, tP, al, wp, dh, id32);
                    t3_ts<true>(nomma || nosmall, tP, ah, wp + 640, dh, id32);
                    t3_ts<true>(nomma || nosmall, tP, al + 16, wp + 320, dh, id32);
                    t3_ts<true>(nomma || nosmall, tP, ah + 16, wp + 960, dh, id32);
                    t3_ts<true>(nomma || nosmall, tP, ah, wp, dh, id32);
                  } else {
                    if (c == 0) t3_ts<false>(nomma || nosmall, tP, ah, wp, dh, id32); else t3_ts<true>(nomma || nosmall, tP, ah, wp, dh, id32);
                  }
                  t3_ts<true>(nomma || nosmall, tP, ah + 16, wp + 320, dh, id32);
                }
                release_chunk();
              }
              tc::mma_commit(&bars->d_full[x]);
            } else {
              // ---- phase 11: reverse of lin0, N = 32 (PE index k in row k + 5), two K = 64 chunks, accumulates
              //      onto the skip layer's contribution in DPE ----
              wait_a(x);
              const uint32_t dh = dh32;
              const uint32_t tP = tbase + T3_DPE(x);
#pragma unroll 1
              for (int c = 0; c < 2; ++c) {
                const uint32_t w0 = (uint32_t)d32 | next_chunk();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint32_t ah = tA + (c * 4 + ks) * 16, al = ah + 8;
                  if (!fast) {
                    t3_ts<true>(nomma || nosmall, tP, al, w0 + ks * 64, dh, id32);
                    t3_ts<true>(nomma || nosmall, tP, ah, w0 + 256 + ks * 64, dh, id32);
                  }
                  t3_ts<true>(nomma || nosmall, tP, ah, w0 + ks * 64, dh, id32);
                }
                release_chunk();
              }
              tc::mma_commit(&bars->tile_done[x]);
            }
          }
        }
      }
      T3_PROF_FLUSH(true, 0, 4)
    }
  } else if (warp >= T3_EPI_WARPS + 2) {
    // =============================== helper warps: staging and the final gradient ===============================
    const int r = (warp & 3) * 32 + lane;          // my row = my TMEM lane (a warp reaches lane quarter warp % 4)
    const uint32_t tl = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    int32_t* rows_cta = rows_all + (size_t)blockIdx.x * T3_NSTAGE * 32 * 128;
    auto put_k = [&](uint8_t* base, int k, float v) {
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const uint32_t off = (uint32_t)(k >> 3) * 2048u + r * 16 + (k & 7) * 2;
      *reinterpret_cast<__half*>(base + off) = h;
      *reinterpret_cast<__half*>(base + 8192 + off) = l;
    };
    // features (4 levels x 7, trilinear) and positional encoding of tile `it` -> smem A operands of buffer it % 4
    auto stage = [&](int64_t it) {
      const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
      const int buf = (int)(it & (T3_NSTAGE - 1));
      uint8_t* afeat = smem + S3_STAGE + buf * S3_STAGE_BYTES;
      uint8_t* ape = afeat + 16384;
      float px, py, pz;
      load_point(tile * 128 + r, px, py, pz);
      // all 32 index loads first, then the voxel rows level by level (batched, branch-free); the row ids are parked
      // for the final gradient of this tile, which then starts with the row fetch
      int32_t rows[4][8];
#pragma unroll
      for (int lv = 0; lv < 4; ++lv) {
        if (lv < sc.n_levels && !T3_DBG(64)) {
          sparse_rows(sc, lv, px, py, pz, rows[lv]);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) rows[lv][c] = -1;
        }
      }
      if (GRAD) {
        int32_t* rb = rows_cta + (size_t)buf * 32 * 128 + r;
#pragma unroll
        for (int lv = 0; lv < 4; ++lv)
#pragma unroll
          for (int c = 0; c < 8; ++c) rb[(lv * 8 + c) * 128] = rows[lv][c];
      }
#pragma unroll
      for (int lv = 0; lv < 4; ++lv) {
        float f7[7];
        sparse_accum<0>(sc, lv, px, py, pz, rows[lv], nullptr, f7);
#pragma unroll
        for (int c = 0; c < 7; ++c) put_k(afeat, lv * 7 + c, f7[c]);
      }
      put_k(afeat, 28, 1.0f);
#pragma unroll
      for (int k = 29; k < 32; ++k) put_k(afeat, k, 0.f);
      const float xs[3] = {px * net.scale, py * net.scale, pz * net.scale};
#pragma unroll
      for (int d = 0; d < 3; ++d) put_k(ape, d, xs[d]);
      float fr = 1.0f;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          float sn = 0.f, cs = 0.f;
          if (f < net.multires) sincosf(xs[d] * fr, &sn, &cs);
          put_k(ape, 3 + 6 * f + d, sn);
          put_k(ape, 3 + 6 * f + 3 + d, cs);
        }
        fr *= 2.0f;
      }
      put_k(ape, 27, 1.0f);
#pragma unroll
      for (int k = 28; k < 32; ++k) put_k(ape, k, 0.f);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->stage_ready[buf]);
    };
    // d sdf / d x of tile `it` = scale * (d PE / d x)^T g_pe + sum over levels (d feat / d x)^T g_feat, with g_feat
    // and g_pe read from the tile's persistent TMEM accumulators
    auto finish = [&](int64_t it) {
      const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
      const int x = (int)(it & 1);
      const uint8_t* ape = smem + S3_STAGE + (it & (T3_NSTAGE - 1)) * S3_STAGE_BYTES + 16384;
      uint32_t gfw[32], gpw[32];
      tc::tmem_ld32(tl + T3_DFEAT(x), gfw);
      tc::tmem_ld32(tl + T3_DPE(x), gpw);
      tc::tmem_wait_ld();
      // the accumulators are in registers: the tensor pipe may reuse them (tile it + 2) while the gather below runs
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->finish_done[x]);
      float px, py, pz;
      const int64_t id = load_point(tile * 128 + r, px, py, pz);
      float acc[3] = {0.f, 0.f, 0.f};
      const int32_t* rb = rows_cta + (size_t)(it & (T3_NSTAGE - 1)) * 32 * 128 + r;
#pragma unroll
      for (int lv = 0; lv < 4; ++lv) {
        int32_t rows[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) rows[c] = rb[(lv * 8 + c) * 128];
        float g7[7], o3[3];
#pragma unroll
        for (int c = 0; c < 7; ++c) g7[c] = __uint_as_float(gfw[lv * 7 + c]) + sw6[128 + lv * 7 + c] * net.inv_scale;
        sparse_accum<1>(sc, lv, px, py, pz, rows, g7, o3);
        acc[0] += o3[0]; acc[1] += o3[1]; acc[2] += o3[2];
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float gx = __uint_as_float(gpw[T3_PE_SHIFT + d]);
        float fr = 1.0f;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          if (f < net.multires) {
            const float sn = t3_get_k(ape, r, 3 + 6 * f + d), cs = t3_get_k(ape, r, 3 + 6 * f + 3 + d);
            gx += fr * (__uint_as_float(gpw[T3_PE_SHIFT + 3 + 6 * f + d]) * cs -
                        __uint_as_float(gpw[T3_PE_SHIFT + 3 + 6 * f + 3 + d]) * sn);
          }
          fr *= 2.0f;
        }
        gx = fmaf(gx, net.scale, acc[d]);
        if (id >= 0) grad_out[id * 3 + d] = gx;
      }
    };
    // staging runs a whole pair ahead of the tensor pipe: buffers it % 4 hold tiles it, it+1 (in flight) and it+2, it+3
    // (prepared); tile it's buffer is handed to tile it+4 as soon as tile it is through
    T3_PROF_DECL
    for (int64_t it = 0; it < T3_NSTAGE && it < my_tiles; ++it) stage(it);
    for (int64_t it = 0; it < my_tiles; ++it) {
      {
        T3_PROF_BEGIN
        tc::mbar_wait(&bars->tile_done[it & 1], (uint32_t)((it >> 1) & 1));
        T3_PROF_END(1)
      }
      if (GRAD) {
        tc::tc_fence_after();
        if (!T3_DBG(256)) {
          T3_PROF_BEGIN
          finish(it);
          T3_PROF_END(2)
        } else {
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars->finish_done[it & 1]);
        }
      }
      if (it + T3_NSTAGE < my_tiles) {
        T3_PROF_BEGIN
        stage(it + T3_NSTAGE);
        T3_PROF_END(0)
      }
    }
    T3_PROF_FLUSH(warp == T3_EPI_WARPS + 2 && lane == 0, 7, 3)
  } else {
    // =============================== weight loader ===============================
    if (lane == 0) {
      T3_PROF_DECL
      int slot = 0;
      uint32_t par = 1;          // parity of the previous use of this slot (first round: nothing to wait for)
      int64_t issued = 0;
      for (int64_t j = 0; j < n_pairs; ++j) {
        for (int p = 0; p < NPH; ++p) {
          // chunk range of phase p in the stream: lin0 | 5 x (2 feature halves + 4 hidden) | 5 x 4 reverse | 2 (lin0 reverse)
          const int c0 = p == 0 ? 0 : (p < 6 ? 1 + 6 * (p - 1) : (p < 11 ? 31 + 4 * (p - 6) : 51));
          const int nc = p == 0 ? 1 : (p < 6 ? 6 : (p < 11 ? 4 : 2));
          for (int x = 0; x < 2; ++x) {
            if (2 * j + x >= my_tiles) continue;
            for (int c = 0; c < nc; ++c) {
              const int cid = c0 + c;
              if (issued >= T3_NSLOT) {
                T3_PROF_BEGIN
                tc::mbar_wait(&bars->w_empty[slot], par);
                T3_PROF_END(0)
              }
              if (T3_DBG(32)) {
                tc::mbar_arrive(&bars->w_full[slot]);
                ++issued;
                slot = (slot + 1 == T3_NSLOT) ? 0 : slot + 1;
                par ^= (slot == 0);
                continue;
              }
              tc::mbar_arrive_expect_tx(&bars->w_full[slot], stream.bytes[cid]);
              tc::bulk_g2s(smem + S3_RING + slot * T3_SLOT_BYTES, wblob + stream.off[cid], stream.bytes[cid],
                           &bars->w_full[slot]);
              ++issued;
              slot = (slot + 1 == T3_NSLOT) ? 0 : slot + 1;
              par ^= (slot == 0);
            }
          }
        }
      }
      T3_PROF_FLUSH(true, 11, 1)
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == T3_EPI_WARPS) tc::tmem_dealloc<512>(tbase);
}

// ---------------------------------------------------------------------------------------------
// host: the fp16 hi|lo weight stream, in the order the issuer consumes it
//   forward : lin0 (K = 27 + bias column), then per layer lin1..lin5 the two K = 16 feature | bias half chunks
//             (independent of the previous layer, issued first) followed by the four K = 32 hidden chunks;
//   reverse : lin5..lin1 as B[n = input index (160 rows)][k = output index], four K = 32 chunks each; lin0 as
//             B[n = PE index + 5 (32 rows)][k], two K = 64 chunks (the shift lines the PE index up with rows 96.. of
//             the skip layer, which accumulate into the same TMEM columns).
// Chunk image = the K-major no-swizzle canonical smem layout: element (n, kk) at (kk >> 3) * rows * 8 + n * 8 + (kk & 7),
// hi half first, lo half after it.
// ---------------------------------------------------------------------------------------------
static inline uint16_t t3_f2h(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
static inline float t3_h2f(uint16_t b) {
  __half h;
  memcpy(&h, &b, 2);
  return __half2float(h);
}

int surf_build_tc_weights(const std::vector<std::vector<float>>& W, const surf_net_inputs* in, surf_net* net,
                          cudaStream_t st, int (*dev_alloc)(surf_net*, void**, size_t)) {
  T1Stream& S = net->tc_stream;
  memset(&S, 0, sizeof(S));
  std::vector<uint16_t> blob;
  int nc = 0;
  auto add_chunk = [&](int rows, int K) {
    const size_t half = (size_t)rows * K;           // halves
    S.off[nc] = (uint32_t)(blob.size() * 2);
    S.bytes[nc] = (uint32_t)(half * 2 * 2);
    blob.resize(blob.size() + half * 2, 0);
    return blob.size() - half * 2;
  };
  auto put = [&](size_t base, int rows, int K, int n, int kk, float v) {
    const uint16_t hi = t3_f2h(v);
    const uint16_t lo = t3_f2h(v - t3_h2f(hi));
    const size_t off = (size_t)(kk >> 3) * rows * 8 + (size_t)n * 8 + (kk & 7);
    blob[base + off] = hi;
    blob[base + (size_t)rows * K + off] = lo;
  };
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
    const int order[6] = {4, 5, 0, 1, 2, 3};        // feature | bias halves first
    for (int ci = 0; ci < 6; ++ci) {
      const int c = order[ci];
      const int K = c < 4 ? 32 : 16;
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
  {
    const int O = in->out_dim[0], I = in->in_dim[0];
    for (int c = 0; c < 2; ++c) {
      const size_t b = add_chunk(32, 64);
      for (int kk = 0; kk < 64; ++kk) {
        const int k = c * 64 + kk;
        if (k >= O) continue;
        for (int n = 0; n < I && n + T3_PE_SHIFT < 32; ++n) put(b, 32, 64, n + T3_PE_SHIFT, kk, W[0][(size_t)k * I + n]);
      }
      nc++;
    }
  }
  S.n_all = nc;
  if (S.n_fwd != 31 || S.n_all != 53) {
    surf_set_error("tensor-core weight stream: unexpected chunk count %d / %d", S.n_fwd, S.n_all);
    return -1;
  }
  void* p = nullptr;
  int rc = dev_alloc(net, &p, blob.size() * 2);
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice, st));
  SURF_CUDA(cudaStreamSynchronize(st));
  net->tc_blob = (const uint8_t*)p;
  // softplus' code scratch: per CTA and tile parity 5 layers x 4 x 512 threads x 16 B (+ the sign words)
  rc = dev_alloc(net, &p, (size_t)net->n_sm * 2 * T3_SCRATCH_U4 * sizeof(uint4));
  if (rc) return rc;
  net->tc_scratch = p;
  // voxel-row ids of the staged tiles (stage -> final gradient): per CTA T3_NSTAGE x [32 ids][128 points]
  rc = dev_alloc(net, &p, (size_t)net->n_sm * T3_NSTAGE * 32 * 128 * sizeof(int32_t));
  if (rc) return rc;
  net->tc_rows = (int32_t*)p;
  return 0;
}

#ifdef T3_DEBUG
static int g_t3_debug_flags = 0;
extern "C" void surf_debug_flags(int f) { g_t3_debug_flags = f; }
extern "C" void surf_debug_prof_read(unsigned long long* h_out) {      // read and reset
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h_out, g_t3_prof, sizeof(unsigned long long) * 16);
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_t3_prof, z, sizeof(z));
}
#endif

int launch_sdf_tc(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_sdf, float* d_grad,
                  bool negate, bool fast, cudaStream_t st) {
  if (src.n <= 0) return 0;
  int rc = surf_ensure_dyn_smem((const void*)k_sdf_pp<true>, S3_TOTAL);
  if (rc) return rc;
  rc = surf_ensure_dyn_smem((const void*)k_sdf_pp<false>, S3_TOTAL);
  if (rc) return rc;
  const int64_t tiles = (src.n + 127) / 128;
  // two tiles per CTA pass: below 2 x SMs tiles, fewer CTAs with a full pair each beat more CTAs with a lone tile
  int64_t grid = (tiles + 1) / 2;
  if (grid > n->n_sm) grid = n->n_sm;
  int flags = (negate ? 1 : 0) | (fast ? 2 : 0);
#ifdef T3_DEBUG
  flags |= g_t3_debug_flags;
#endif
  surf_time_begin(d_grad ? 0 : 1, st);
  if (d_grad) {
    k_sdf_pp<true><<<(int)grid, T3_THREADS, S3_TOTAL, st>>>(s->dev, n->dev, src, n->tc_blob, n->tc_stream, d_sdf, d_grad,
                                                           (uint4*)n->tc_scratch, n->tc_rows, flags);
  } else {
    k_sdf_pp<false><<<(int)grid, T3_THREADS, S3_TOTAL, st>>>(s->dev, n->dev, src, n->tc_blob, n->tc_stream, d_sdf, nullptr,
                                                            (uint4*)n->tc_scratch, n->tc_rows, flags);
  }
  surf_time_end(d_grad ? 0 : 1, st);
  SURF_LAUNCH_CHECK();
  return 0;
}
