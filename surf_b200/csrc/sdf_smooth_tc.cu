// Tensor-core edition of the second-order term of SDFNetworkSparse.gradient (sdf_network.py:129-152):
//   smooth = d/dx [ sum_j d sdf / d x_j ] = Hessian(sdf) . (1,1,1)        (the math: see sdf_smooth.cu)
// Forward-over-reverse needs every Linear layer four times — primal and tangent, forward and reverse — with the SAME
// weights, so a 128-point tile runs each layer as two M = 128 tcgen05 GEMMs (primal stream, tangent stream) that share
// the B operand; thread r of a lane quarter owns point r in both accumulators, so the coupling between the streams
// (hd = s'(z) zd,  deltad = gad s' + ga s'' zd) is register arithmetic.
//
// Per CTA (one per SM, persistent over tiles): 8 epilogue warps (warp (q, hf): TMEM lane quarter q, column half hf)
// + one warp whose elected lane streams the weights (cp.async.bulk, one layer ahead of the epilogue) and issues the
// MMAs.  fp16 hi/lo split, 3 MMAs per product (fp32-grade, like sdf_tc2.cu); SURF_MLP_TC_FAST issues one.
//   TMEM (448 of 512 columns): primal A operand hi 64 | lo 64 (TS form), accumulators D0 (primal) 128 | D1 (tangent)
//         128, feature-gradient accumulators F0 | F1 32 each (accumulated over the reverse layers by the tensor core)
//   smem: weights of the current layer 80 KB (hidden 128x128 hi|lo, feature / bias block 128x32 hi|lo), tangent A
//         operand 64 KB (SS form), feature operands (constant over the layers, bias folded in as a ones column) 32 KB,
//         positional encoding fp32 13.5 KB
//   L2 scratch (per CTA): (s'(z_l), s''(z_l) zd_l) of layers 0..4, 640 KB, written and read by the same thread
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "smooth_common.cuh"
#include "surf_internal.cuh"
#include "tc_common.cuh"
#include "tc_layer_common.cuh"

#define ST_EPI_WARPS 8
#define ST_THREADS ((ST_EPI_WARPS + 1) * 32)
// TMEM columns
#define ST_AP_HI 0u
#define ST_AP_LO 64u
#define ST_D0 128u
#define ST_D1 256u
#define ST_F0 384u
#define ST_F1 416u
// shared memory map
#define ST_SM_W 0
#define ST_SM_AT ST_W_BYTES                  // tangent A operand: hi 32 KB | lo 32 KB
#define ST_SM_AFP (ST_SM_AT + 65536)         // primal feature operand K = 32: hi 8 KB | lo 8 KB
#define ST_SM_AFT (ST_SM_AFP + 16384)        // tangent feature operand
#define ST_SM_PE (ST_SM_AFT + 16384)         // float PE[27][128]
#define ST_SM_PART (ST_SM_PE + 27 * 128 * 4) // float part[128][12]: what column half 1 hands to column half 0
#define ST_SM_BAR (ST_SM_PART + 128 * 12 * 4)
#define ST_SM_TOTAL (ST_SM_BAR + 64)
#define ST_SCRATCH_FLOAT2 (5 * 128 * 128)    // per CTA

static_assert(ST_SM_TOTAL <= 227 * 1024, "shared memory of k_sdf_smooth_tc");

struct StBars {
  uint64_t a_ready;   // the A operands of the next step are in place: one arrival per epilogue warp
  uint64_t d_full;    // the MMAs of a step are done (tcgen05.commit)
  uint64_t w_full;    // the weights of a step have landed (expect_tx)
  uint32_t tmem_base;
};

// contribution of the gradient (g, gd) w.r.t. PE input j of my point to d/dx (g1) and to the second-order term (s2)
// (per unit of the scaled coordinate X = scale * x; the caller multiplies by scale)
__device__ __forceinline__ void st_pe_accum(const float* PE, int r, int j, float scale, float g, float gd, float (&g1)[3],
                                            float (&s2)[3]) {
  if (j < 3) {
    g1[j] += g;
    s2[j] += gd;
    return;
  }
  const int t = j - 3, f = t / 6, rem = t % 6, d = rem % 3;
  const float fr = (float)(1 << f);
  const float sn = PE[(3 + 6 * f + d) * 128 + r], cs = PE[(3 + 6 * f + 3 + d) * 128 + r];
  if (rem < 3) {      // sin(fr X): d/dX = fr cos, d2/dX2 = -fr^2 sin
    g1[d] = fmaf(fr * g, cs, g1[d]);
    s2[d] = fmaf(fr * gd, cs, s2[d]);
    s2[d] = fmaf(-fr * fr * scale * g, sn, s2[d]);
  } else {            // cos(fr X): d/dX = -fr sin, d2/dX2 = -fr^2 cos
    g1[d] = fmaf(-fr * g, sn, g1[d]);
    s2[d] = fmaf(-fr * gd, sn, s2[d]);
    s2[d] = fmaf(-fr * fr * scale * g, cs, s2[d]);
  }
}

__global__ void __launch_bounds__(ST_THREADS, 1)
k_sdf_smooth_tc(const DevScene sc, const DevNet net, const uint8_t* __restrict__ wblob, const float* __restrict__ pts,
                const uint8_t* __restrict__ flags, int64_t n, float* __restrict__ grad_out, float* __restrict__ smooth_out,
                float2* __restrict__ scratch_all, int fast_i) {
  const bool fast = fast_i != 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  StBars* bars = reinterpret_cast<StBars*>(smem + ST_SM_BAR);
  float* PE = reinterpret_cast<float*>(smem + ST_SM_PE);
  float* PART = reinterpret_cast<float*>(smem + ST_SM_PART);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // feature operands: zero, ones column (bias row of the weights) at k = 28 of the primal operand
  for (int i = tid; i < 32768 / 16; i += ST_THREADS) reinterpret_cast<uint4*>(smem + ST_SM_AFP)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == ST_EPI_WARPS) tc::tmem_alloc<512>(&bars->tmem_base);
  if (tid == 0) {
    tc::mbar_init(&bars->a_ready, ST_EPI_WARPS);
    tc::mbar_init(&bars->d_full, 1);
    tc::mbar_init(&bars->w_full, 1);
    tc::mbar_fence_init();
  }
  __syncthreads();
  if (tid < ST_ROWS) *reinterpret_cast<__half*>(smem + ST_SM_AFP + 3 * 2048 + tid * 16 + 4 * 2) = __float2half_rn(1.0f);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = bars->tmem_base;
  const int64_t n_tiles = (n + ST_ROWS - 1) / ST_ROWS;
  const float scale = net.scale;

  if (warp < ST_EPI_WARPS) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t tl = tbase + ((uint32_t)(q * 32) << 16);
    const uint32_t t_hi = tl + ST_AP_HI, t_lo = tl + ST_AP_LO;
    float2* scratch = scratch_all + (size_t)blockIdx.x * ST_SCRATCH_FLOAT2;
    const int pair_bar = 1 + q;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    auto signal = [&]() {
      tc::tmem_wait_st();
      tc::fence_proxy_async();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->a_ready);
    };
    uint32_t ph = 0;
    auto wait_d = [&]() {
      tc::mbar_wait(&bars->d_full, ph & 1);
      ph++;
      tc::tc_fence_after();
    };
    const float c6s = net.inv_scale;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t i = tile * ST_ROWS + r;
      const bool inb = i < n;
      const float px = inb ? pts[i * 3] : 0.f, py = inb ? pts[i * 3 + 1] : 0.f, pz = inb ? pts[i * 3 + 2] : 0.f;
      // ---- inputs: positional encoding + tangent (column half 0), two feature levels each ----
      if (hf == 0) {
        float pe[32], ped[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) { pe[k] = 0.f; ped[k] = 0.f; }
        const float x[3] = {px, py, pz};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float X = x[d] * scale;
          pe[d] = X;
          ped[d] = scale;
          float fr = 1.0f;
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            float sn, cs;
            sincosf(X * fr, &sn, &cs);
            pe[3 + 6 * f + d] = sn;      ped[3 + 6 * f + d] = fr * scale * cs;
            pe[3 + 6 * f + 3 + d] = cs;  ped[3 + 6 * f + 3 + d] = -fr * scale * sn;
            fr *= 2.0f;
          }
        }
#pragma unroll
        for (int k = 0; k < 27; ++k) PE[k * 128 + r] = pe[k];
        pe[27] = 1.0f;                 // bias row of lin0
        {
          float a[16], b[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) { a[k] = pe[k]; b[k] = pe[16 + k]; }
          st_store_tmem(t_hi, t_lo, 0, a);
          st_store_tmem(t_hi, t_lo, 16, b);
#pragma unroll
          for (int k = 0; k < 16; ++k) { a[k] = ped[k]; b[k] = ped[16 + k]; }
          st_store_smem(smem + ST_SM_AT, 32768, r, 0, a);
          st_store_smem(smem + ST_SM_AT, 32768, r, 16, b);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int lv = hf * 2 + u;
        float f7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, fd7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (inb && lv < sc.n_levels) sparse_value_tangent_batched(sc, lv, px, py, pz, f7, fd7);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          st_put_half(smem + ST_SM_AFP, 8192, r, lv * 7 + c, f7[c]);
          st_put_half(smem + ST_SM_AFT, 8192, r, lv * 7 + c, fd7[c]);
        }
      }
      pair_sync();            // PE of my point is visible to the other column half
      signal();
      // ---- forward with tangent: lin0 .. lin5 ----
#pragma unroll 1
      for (int l = 0; l < 6; ++l) {
        wait_d();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cb = hf * 64 + c * 16;
          uint32_t z[16], zd[16];
          tc::tmem_ld16(tl + ST_D0 + cb, z);
          tc::tmem_ld16(tl + ST_D1 + cb, zd);
          tc::tmem_wait_ld();
          float a[16], b[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = cb + j;
            float h, d1, d2;
            st_softplus(__uint_as_float(z[j]), h, d1, d2);
            const float t = __uint_as_float(zd[j]);
            if (l < 5) {
              a[j] = h;
              b[j] = d1 * t;
              scratch[(size_t)(l * 128 + col) * 128 + r] = make_float2(d1, d2 * t);
            } else {         // delta_5 = ga_6 s'(z_5), its tangent ga_6 s''(z_5) zd_5 (gad_6 = 0)
              const float w = net.w6[col] * c6s;
              a[j] = w * d1;
              b[j] = w * d2 * t;
            }
          }
          if (l == 2 && hf == 1 && c >= 2) {     // the skip layer's input: columns 101..127 are the positional encoding
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = 64 + c * 16 + j - 101;        // compile-time
              if (k >= 0) {
                a[j] = PE[k * 128 + r];
                if (k < 3) {
                  b[j] = scale;
                } else {
                  const int t6 = (k - 3) % 6, f = (k - 3) / 6, d = t6 % 3;
                  const float fr = (float)(1 << f);
                  b[j] = t6 < 3 ? fr * scale * PE[(3 + 6 * f + 3 + d) * 128 + r] : -fr * scale * PE[(3 + 6 * f + d) * 128 + r];
                }
              }
            }
          }
          st_store_tmem(t_hi, t_lo, cb, a);
          st_store_smem(smem + ST_SM_AT, 32768, r, cb, b);
        }
        signal();
      }
      // ---- reverse with tangent: lin5 .. lin1 ----
      float g1[3] = {0.f, 0.f, 0.f}, s2[3] = {0.f, 0.f, 0.f};       // PE part of d/dx and of the second-order term
#pragma unroll 1
      for (int l = 5; l >= 1; --l) {
        const bool pe_layer = (l == 3) && (hf == 1);
        // (s', s'' zd) of layer l-1, one chunk of 16 columns ahead of its use: the first chunk is fetched before the
        // accumulator wait, the loads never sit between a TMEM read and its use
        const float2* sl = scratch + (size_t)((l - 1) * 128 + hf * 64) * 128 + r;
        float2 dn[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) dn[j] = sl[(size_t)j * 128];
        wait_d();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cb = hf * 64 + c * 16;
          float2 dc[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) dc[j] = dn[j];
          if (c < 3) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dn[j] = sl[(size_t)((c + 1) * 16 + j) * 128];
          }
          uint32_t ga[16], gad[16];
          tc::tmem_ld16(tl + ST_D0 + cb, ga);
          tc::tmem_ld16(tl + ST_D1 + cb, gad);
          tc::tmem_wait_ld();
          float a[16], b[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float g = __uint_as_float(ga[j]), gd = __uint_as_float(gad[j]);
            if (c >= 2 && 64 + c * 16 + j >= 101 && pe_layer) {       // gradient w.r.t. the PE part of the skip layer's input
              st_pe_accum(PE, r, 64 + c * 16 + j - 101, scale, g, gd, g1, s2);
              a[j] = 0.f;
              b[j] = 0.f;
            } else {
              a[j] = g * dc[j].x;
              b[j] = fmaf(gd, dc[j].x, g * dc[j].y);
            }
          }
          st_store_tmem(t_hi, t_lo, cb, a);
          st_store_smem(smem + ST_SM_AT, 32768, r, cb, b);
        }
        signal();
      }
      // ---- lin0's input gradient = PE gradient; feature gradients from the F accumulators; d/dx ----
      wait_d();
      float o3[3] = {0.f, 0.f, 0.f}, sm3[3] = {0.f, 0.f, 0.f};     // feature part of the gradient / second-order term
      {
        float gf[14], gfd[14];
        if (hf == 0) {
          uint32_t a[16], b[16];
          tc::tmem_ld16(tl + ST_F0, a);
          tc::tmem_ld16(tl + ST_F1, b);
          tc::tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 14; ++k) { gf[k] = __uint_as_float(a[k]); gfd[k] = __uint_as_float(b[k]); }
        } else {
          uint32_t a8[8], b8[8], a[16], b[16];
          tc::tmem_ld8(tl + ST_F0 + 8, a8);
          tc::tmem_ld8(tl + ST_F1 + 8, b8);
          tc::tmem_ld16(tl + ST_F0 + 16, a);
          tc::tmem_ld16(tl + ST_F1 + 16, b);
          tc::tmem_wait_ld();
          gf[0] = __uint_as_float(a8[6]); gf[1] = __uint_as_float(a8[7]);
          gfd[0] = __uint_as_float(b8[6]); gfd[1] = __uint_as_float(b8[7]);
#pragma unroll
          for (int k = 0; k < 12; ++k) { gf[2 + k] = __uint_as_float(a[k]); gfd[2 + k] = __uint_as_float(b[k]); }
        }
#pragma unroll
        for (int k = 0; k < 14; ++k) gf[k] = fmaf(net.w6[128 + hf * 14 + k], c6s, gf[k]);    // lin6 sees the features too
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int lv = hf * 2 + u;
          if (inb && lv < sc.n_levels) {
            float g3[3], od3[3], om3[3];
            // J^T g and J^T gd (scaled 1/vs), the mixed second derivatives contracted with g (1/vs^2)
            sparse_back_fused(sc, lv, px, py, pz, &gf[u * 7], &gfd[u * 7], g3, od3, om3);
            const float inv = 1.0f / sc.voxel[lv];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              o3[d] += g3[d];
              sm3[d] += fmaf(om3[d], inv * inv, od3[d]);
            }
          }
        }
      }
      if (hf == 0) {
        uint32_t ga[16], gad[16], gb[16], gbd[16];
        tc::tmem_ld16(tl + ST_D0, ga);
        tc::tmem_ld16(tl + ST_D1, gad);
        tc::tmem_ld16(tl + ST_D0 + 16, gb);
        tc::tmem_ld16(tl + ST_D1 + 16, gbd);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) st_pe_accum(PE, r, j, scale, __uint_as_float(ga[j]), __uint_as_float(gad[j]), g1, s2);
#pragma unroll
        for (int j = 16; j < 27; ++j) st_pe_accum(PE, r, j, scale, __uint_as_float(gb[j - 16]), __uint_as_float(gbd[j - 16]), g1, s2);
      } else {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          PART[r * 12 + d] = g1[d]; PART[r * 12 + 3 + d] = s2[d]; PART[r * 12 + 6 + d] = o3[d]; PART[r * 12 + 9 + d] = sm3[d];
        }
      }
      tc::tc_fence_before();
      pair_sync();
      if (hf == 0 && inb) {
        const bool on = flags == nullptr || ((flags[i] >> 1) & 1);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float gx = fmaf(g1[d] + PART[r * 12 + d], scale, o3[d] + PART[r * 12 + 6 + d]);
          const float sx = fmaf(s2[d] + PART[r * 12 + 3 + d], scale, sm3[d] + PART[r * 12 + 9 + d]);
          smooth_out[i * 3 + d] = on ? sx : 0.f;
          if (grad_out) grad_out[i * 3 + d] = on ? gx : 0.f;
        }
      }
      pair_sync();            // PART / PE are free for the next tile
    }
  } else {
    // =============================== weight stream + MMA issue ===============================
    if (tc::elect_one()) {
      const uint32_t w_a = tc::smem_u32(smem + ST_SM_W);
      const uint32_t at_hi = tc::smem_u32(smem + ST_SM_AT), at_lo = at_hi + 32768u;
      const uint32_t afp_hi = tc::smem_u32(smem + ST_SM_AFP), afp_lo = afp_hi + 8192u;
      const uint32_t aft_hi = tc::smem_u32(smem + ST_SM_AFT), aft_lo = aft_hi + 8192u;
      const uint32_t ap_hi = tbase + ST_AP_HI, ap_lo = tbase + ST_AP_LO;
      const uint32_t D0 = tbase + ST_D0, D1 = tbase + ST_D1, F0 = tbase + ST_F0, F1 = tbase + ST_F1;
      auto load = [&](int step) {
        const uint32_t bytes = st_step_bytes(step);
        tc::mbar_arrive_expect_tx(&bars->w_full, bytes);
        for (uint32_t o = 0; o < bytes; o += 16384u)
          tc::bulk_g2s(smem + ST_SM_W + o, wblob + st_step_off(step) + o, 16384u, &bars->w_full);
      };
      uint32_t g = 0;        // steps issued so far: parity of all three barriers
      if ((int64_t)blockIdx.x < n_tiles) load(0);
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int step = 0; step < ST_STEPS; ++step, ++g) {
          tc::mbar_wait(&bars->a_ready, g & 1);
          tc::mbar_wait(&bars->w_full, g & 1);
          tc::tc_fence_after();
          if (step == 0) {                       // lin0 forward: K = 32 (27 PE + ones)
            st_gemm<128, 2, true>(D0, ap_hi, ap_lo, w_a, 8192u, false, fast);
            st_gemm<128, 2, false>(D1, at_hi, at_lo, w_a, 8192u, false, fast);
          } else if (step < 6) {                 // lin1..lin5 forward: hidden K = 128 + feature / bias K = 32
            st_gemm<128, 8, true>(D0, ap_hi, ap_lo, w_a, ST_W_HID_LO, false, fast);
            st_gemm<128, 2, false>(D0, afp_hi, afp_lo, w_a + ST_W_FEAT, 8192u, true, fast);
            st_gemm<128, 8, false>(D1, at_hi, at_lo, w_a, ST_W_HID_LO, false, fast);
            st_gemm<128, 2, false>(D1, aft_hi, aft_lo, w_a + ST_W_FEAT, 8192u, true, fast);
          } else if (step < ST_STEPS - 1) {      // lin5..lin1 reverse: hidden columns -> D, feature columns -> F (+=)
            const bool acc = step != 6;
            st_gemm<128, 8, true>(D0, ap_hi, ap_lo, w_a, ST_W_HID_LO, false, fast);
            st_gemm<32, 8, true>(F0, ap_hi, ap_lo, w_a + ST_W_FEAT, 8192u, acc, fast);
            st_gemm<128, 8, false>(D1, at_hi, at_lo, w_a, ST_W_HID_LO, false, fast);
            st_gemm<32, 8, false>(F1, at_hi, at_lo, w_a + ST_W_FEAT, 8192u, acc, fast);
          } else {                               // lin0 reverse: 27 PE columns
            st_gemm<32, 8, true>(D0, ap_hi, ap_lo, w_a, 8192u, false, fast);
            st_gemm<32, 8, false>(D1, at_hi, at_lo, w_a, 8192u, false, fast);
          }
          tc::mma_commit(&bars->d_full);
          // the weight buffer is free once these MMAs are done; fetch the next step's while the epilogue runs
          const bool more = step + 1 < ST_STEPS || tile + gridDim.x < n_tiles;
          if (more) {
            tc::mbar_wait(&bars->d_full, g & 1);
            load(step + 1 < ST_STEPS ? step + 1 : 0);
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == ST_EPI_WARPS) tc::tmem_dealloc<512>(tbase);
}

// ---------------------------------------------------------------------------------------------
// host: weight blob in the order the kernel streams it
// ---------------------------------------------------------------------------------------------
static inline uint16_t st_f2h(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
static inline float st_h2f(uint16_t b) {
  __half h;
  memcpy(&h, &b, 2);
  return __half2float(h);
}
// canonical K-major operand with N rows, K columns: hi at off, lo at off + N * K * 2
template <typename F>
static void st_put(std::vector<uint8_t>& blob, size_t off, int N, int K, F value) {
  uint16_t* hi = reinterpret_cast<uint16_t*>(blob.data() + off);
  uint16_t* lo = hi + (size_t)N * K;
  for (int nn = 0; nn < N; ++nn)
    for (int k = 0; k < K; ++k) {
      const float v = value(nn, k);
      const uint16_t h = st_f2h(v);
      const size_t idx = (size_t)(k >> 3) * N * 8 + (size_t)nn * 8 + (k & 7);
      hi[idx] = h;
      lo[idx] = st_f2h(v - st_h2f(h));
    }
}

int surf_build_smooth_tc_weights(const std::vector<std::vector<float>>& W, const surf_net_inputs* in, surf_net* net,
                                 cudaStream_t st, int (*dev_alloc)(surf_net*, void**, size_t)) {
  std::vector<uint8_t> blob(ST_BLOB_BYTES, 0);
  const int pe = in->in_dim[0];      // 27
  {   // step 0: lin0 forward, N = 128 outputs x K = 32 (PE, bias at k = 27)
    const int O = in->out_dim[0], I = in->in_dim[0];
    st_put(blob, st_step_off(0), 128, 32, [&](int o, int k) {
      if (o >= O) return 0.f;
      return k < I ? W[0][(size_t)o * I + k] : (k == pe ? in->h_bias[0][o] : 0.f);
    });
  }
  for (int l = 1; l < 6; ++l) {   // steps 1..5: hidden 128 x 128, then feature / bias block 128 x 32 (bias at k = 28)
    const int O = in->out_dim[l], I = in->in_dim[l];
    const size_t off = st_step_off(l);
    st_put(blob, off, 128, 128, [&](int o, int k) { return o < O ? W[l][(size_t)o * I + k] : 0.f; });
    st_put(blob, off + ST_W_FEAT, 128, 32, [&](int o, int k) {
      if (o >= O) return 0.f;
      return k < 28 ? W[l][(size_t)o * I + 128 + k] : (k == 28 ? in->h_bias[l][o] : 0.f);
    });
  }
  for (int l = 5; l >= 1; --l) {  // steps 6..10: reverse, N = input column, K = output row
    const int O = in->out_dim[l], I = in->in_dim[l];
    const size_t off = st_step_off(6 + (5 - l));
    st_put(blob, off, 128, 128, [&](int j, int o) { return o < O ? W[l][(size_t)o * I + j] : 0.f; });
    st_put(blob, off + ST_W_FEAT, 32, 128, [&](int j, int o) { return (o < O && j < 28) ? W[l][(size_t)o * I + 128 + j] : 0.f; });
  }
  {   // step 11: lin0 reverse, N = 32 (27 PE columns) x K = 128
    const int O = in->out_dim[0], I = in->in_dim[0];
    st_put(blob, st_step_off(ST_STEPS - 1), 32, 128, [&](int j, int o) { return (o < O && j < I) ? W[0][(size_t)o * I + j] : 0.f; });
  }
  void* p = nullptr;
  int rc = dev_alloc(net, &p, blob.size());
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
  net->smooth_tc_w = (const uint8_t*)p;
  rc = dev_alloc(net, &p, (size_t)net->n_sm * ST_SCRATCH_FLOAT2 * sizeof(float2));
  if (rc) return rc;
  net->smooth_tc_scratch = (float2*)p;
  SURF_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int launch_sdf_smooth_tc(const surf_scene* s, const surf_net* n, const float* d_pts, int64_t n_pts, const uint8_t* d_flags,
                         float* d_grad, float* d_smooth, bool fast, cudaStream_t st) {
  if (n_pts <= 0) return 0;
  SURF_CHECK_ARG(n->smooth_tc_w && n->smooth_tc_scratch, "network without tensor-core second-order weights");
  int rc = surf_ensure_dyn_smem((const void*)k_sdf_smooth_tc, ST_SM_TOTAL);
  if (rc) return rc;
  const int64_t tiles = (n_pts + ST_ROWS - 1) / ST_ROWS;
  const int grid = (int)(tiles < n->n_sm ? tiles : n->n_sm);
  k_sdf_smooth_tc<<<grid, ST_THREADS, ST_SM_TOTAL, st>>>(s->dev, n->dev, n->smooth_tc_w, d_pts, d_flags, n_pts, d_grad,
                                                         d_smooth, n->smooth_tc_scratch, fast ? 1 : 0);
  SURF_LAUNCH_CHECK();
  return 0;
}
