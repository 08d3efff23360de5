// K3b, tensor-core edition: BlendingNetwork.forward (blending_network.py:69-117) on tcgen05.
//
// Tile = 128 (point, view) rows = the 128 TMEM lanes.  The six Linear layers that carry 95 % of the MACs
// (base_fc 57->64->32, vis_fc 32->32->33, vis_fc2.0 32->32, rgb_fc.0 37->16) run as M=128 tcgen05 MMAs: A operand =
// the activations in shared memory (K-major canonical layout, fp16 hi + lo), B operand = the layer's weights,
// resident in shared memory as fp16 hi + lo, accumulator in TMEM; every product hi*hi + lo*hi + hi*lo (fp32-grade).
// The tiny layers (ray_dir_fc 4->16->19, vis_fc2.2 32->1, rgb_fc.2/.4 16->8->1) and the cross-view operations
// (pooling weights, weighted mean / variance, masked softmax) stay on the CUDA cores.  Layers are serial within a
// tile; two CTAs per SM (each < 112 KB smem, 64 TMEM columns) overlap one tile's MMAs with the other's epilogue.
#include <math.h>
#include <string.h>

#include <vector>

#include "lookup_common.cuh"
#include "surf_internal.cuh"
#include "tc_common.cuh"

#define BT_THREADS 256
#define BT_ROWS 128
#define BS 132                      // row stride (floats) of the k-major fp32 buffers

// ---- shared memory map (bytes) ----
// tensor-core weights: [hi | lo] per layer, canonical K-major, (k8 group g, n) at g * N * 16 + n * 16
#define W_BASE0 0                   // N 64, K 64 : 2 x 8192
#define W_BASE1 (W_BASE0 + 16384)   // N 32, K 64 : 2 x 4096
#define W_VIS0 (W_BASE1 + 8192)     // N 32, K 32 : 2 x 2048
#define W_VIS1 (W_VIS0 + 4096)      // N 48, K 32 : 2 x 3072
#define W_V20 (W_VIS1 + 6144)       // N 32, K 32 : 2 x 2048
#define W_RGB0 (W_V20 + 4096)       // N 16, K 48 : 2 x 1536
#define W_TC_BYTES (W_RGB0 + 3072)  // 41984
// fp32 parameters (floats, offsets inside the WF block)
#define F_DIR0_W 0                  // [16][4]
#define F_DIR0_B 64
#define F_DIR1_W 80                 // [19][16]
#define F_DIR1_B 384                // 19 (+1 pad)
#define F_BASE0_B 404               // 64
#define F_BASE1_B 468               // 32
#define F_VIS0_B 500                // 32
#define F_VIS1_B 532                // 33 (+3 pad)
#define F_V20_B 568                 // 32
#define F_V21_W 600                 // 32
#define F_V21_B 632                 // 1 (+3)
#define F_RGB0_B 636                // 16
#define F_RGB1_W 652                // [8][16]
#define F_RGB1_B 780                // 8
#define F_RGB2_W 788                // 8
#define F_RGB2_B 796                // 1 (+3)
#define F_TOTAL 800
#define SB_WF W_TC_BYTES
#define SB_AOP (SB_WF + F_TOTAL * 4)          // 32 KB: [hi 16 KB | lo 16 KB], K <= 64
#define SB_XF (SB_AOP + 32768)                // x fp32 [32][BS]  (first used as h16 of ray_dir_fc)
#define SB_F (SB_XF + 32 * BS * 4)            // [19][BS] fp32: feat19, then x19 in place; later a16 [16][BS]
#define SB_RD (SB_F + 19 * BS * 4)            // [4][BS]
#define SB_RGB (SB_RD + 4 * BS * 4)           // [3][BS] source RGB
#define SB_SC (SB_RGB + 3 * BS * 4)           // scalars: wv, m, vis, logit, e, part : 6 x 128 floats
#define SB_BAR (SB_SC + 6 * 128 * 4)
#define SB_TOTAL (SB_BAR + 64)

struct BtBars {
  uint64_t mma_done;
  uint32_t tmem_base;
};

// ELU; the exponential is the bare MUFU.EX2 (flush-to-zero: exp underflows to 0 exactly where exp(x) - 1 == -1 in
// fp32 anyway) — __expf wraps it in a denormal-range rescue that doubles the instruction count of the hottest
// function of this kernel
__device__ __forceinline__ float bt_elu(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return x > 0.f ? x : e - 1.0f;
}
__device__ __forceinline__ float bt_sigmoid(float x) { return 1.0f / (1.0f + __expf(-x)); }

// write 16 consecutive k values (k0 .. k0+15, k0 % 8 == 0) of row r into the canonical A operand
__device__ __forceinline__ void aop_store16(uint8_t* aop, int r, int k0, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) tc::split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  uint8_t* p = aop + (uint32_t)(k0 >> 3) * 2048u + r * 16;
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  *reinterpret_cast<uint4*>(p + 16384) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  *reinterpret_cast<uint4*>(p + 16384 + 2048) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}
__device__ __forceinline__ void aop_store8(uint8_t* aop, int r, int k0, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) tc::split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  uint8_t* p = aop + (uint32_t)(k0 >> 3) * 2048u + r * 16;
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p + 16384) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// one layer on the tensor core: D[128 x N] = A_op[128 x K] * W^T, issued by one thread
template <int N, int KSTEPS>
__device__ __forceinline__ void bt_issue(uint32_t tD, uint32_t aop_addr, uint32_t w_addr, uint64_t* bar, bool fast) {
  const uint32_t idesc = tc::idesc_f16(128, N, 0);
  const uint64_t da = tc::smem_desc_kmajor(0, 2048, 128), db = tc::smem_desc_kmajor(0, N * 16, 128);
  const uint32_t ah = (uint32_t)(da >> 32), bh = (uint32_t)(db >> 32);
  const uint32_t a0 = (uint32_t)da | (aop_addr >> 4), b0 = (uint32_t)db | (w_addr >> 4);
  constexpr uint32_t A_LO = 16384 >> 4, B_LO = (N * KSTEPS * 16 * 2) >> 4;   // lo halves
  constexpr uint32_t A_KS = 4096 >> 4, B_KS = (N * 32) >> 4;                 // one K step = two 8-wide groups
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    if (ks == 0) tc::mma_ss_w<false>(tD, a0, ah, b0, bh, idesc);
    else tc::mma_ss_w<true>(tD, a0 + ks * A_KS, ah, b0 + ks * B_KS, bh, idesc);
    if (!fast) tc::mma_ss_w<true>(tD, a0 + A_LO + ks * A_KS, ah, b0 + ks * B_KS, bh, idesc);
    if (!fast) tc::mma_ss_w<true>(tD, a0 + ks * A_KS, ah, b0 + B_LO + ks * B_KS, bh, idesc);
  }
  tc::mma_commit(bar);
}

// VT: number of source views at compile time (2, 3, 4: the per-point loops over the views unroll) or 0 = runtime V
template <int VT>
__global__ void __launch_bounds__(BT_THREADS, 2)
k_blend_tc(const uint8_t* __restrict__ wtc, const float* __restrict__ wf32, float s_abs, const float* __restrict__ feat,
           const float* __restrict__ rdiff, const uint8_t* __restrict__ mask, int V_rt, int packed19,
           const int32_t* __restrict__ list, const int32_t* __restrict__ count, int64_t n, float* __restrict__ rgb_out,
           uint8_t* __restrict__ views_out, int fast_i) {
  const bool fast = fast_i != 0;
  const int V = VT ? VT : V_rt;
  extern __shared__ __align__(1024) uint8_t smem[];
  BtBars* bars = reinterpret_cast<BtBars*>(smem + SB_BAR);
  const float* WF = reinterpret_cast<const float*>(smem + SB_WF);
  uint8_t* aop = smem + SB_AOP;
  float* XF = reinterpret_cast<float*>(smem + SB_XF);
  float* F = reinterpret_cast<float*>(smem + SB_F);
  float* RD = reinterpret_cast<float*>(smem + SB_RD);
  float* RGB = reinterpret_cast<float*>(smem + SB_RGB);
  float* s_wv = reinterpret_cast<float*>(smem + SB_SC);
  float* s_m = s_wv + 128;
  float* s_vis = s_m + 128;
  float* s_logit = s_vis + 128;
  float* s_e = s_logit + 128;
  float* s_part = s_e + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane;                     // my row / TMEM lane
  // ---- one-time: weights -> smem, TMEM, barrier ----
  for (int i = tid; i < W_TC_BYTES / 16; i += BT_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(wtc)[i];
  for (int i = tid; i < F_TOTAL; i += BT_THREADS) reinterpret_cast<float*>(smem + SB_WF)[i] = wf32[i];
  if (warp == 0) tc::tmem_alloc<64>(&bars->tmem_base);
  if (tid == 0) {
    tc::mbar_init(&bars->mma_done, 1);
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = bars->tmem_base;
  const uint32_t tl = tbase + ((uint32_t)(q * 32) << 16);
  const uint32_t aop_a = tc::smem_u32(aop), w_a = tc::smem_u32(smem);
  uint32_t phase = 0;

  int64_t n_total = n;
  if (count) {
    const int64_t c = *count;
    n_total = c < n_total ? c : n_total;
  }
  const int ppt = BT_ROWS / V;
  const int rows = ppt * V;
  const int64_t n_tiles = (n_total + ppt - 1) / ppt;

  // wait for the layer's MMAs, then make the accumulator readable
  auto wait_mma = [&]() {
    tc::mbar_wait(&bars->mma_done, phase & 1);
    phase++;
    tc::tc_fence_after();
  };
  // all epilogue writes to the A operand done -> visible to the tensor core, then thread 0 may issue
  auto publish = [&]() {
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  };

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- load records: row r <-> (point tile*ppt + r / V, view r % V) ----
    if (tid < BT_ROWS) {
      const int64_t rec_i = tile * rows + tid;
      const bool ok = (tid < rows) && (rec_i < n_total * V);
      float rec[FEAT_REC];
      float4 rd = make_float4(0.f, 0.f, 0.f, 1.f);
#pragma unroll
      for (int c = 0; c < FEAT_REC; ++c) rec[c] = 0.f;
      if (ok) {
        rd = reinterpret_cast<const float4*>(rdiff)[rec_i];
        if (packed19) {
          const float* f = feat + rec_i * 19;
#pragma unroll
          for (int c = 0; c < 19; ++c) rec[c] = f[c];
          rec[19] = mask[rec_i] ? 1.0f : 0.0f;
        } else {
          const float4* f = reinterpret_cast<const float4*>(feat + rec_i * FEAT_REC);
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            const float4 t = f[c];
            rec[4 * c] = t.x; rec[4 * c + 1] = t.y; rec[4 * c + 2] = t.z; rec[4 * c + 3] = t.w;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 19; ++c) F[c * BS + tid] = rec[c];
      RGB[tid] = rec[0]; RGB[BS + tid] = rec[1]; RGB[2 * BS + tid] = rec[2];
      RD[tid] = rd.x; RD[BS + tid] = rd.y; RD[2 * BS + tid] = rd.z; RD[3 * BS + tid] = rd.w;
      s_m[tid] = rec[19];
      // pooling exponential, correctly rounded (see blend.cu / DESIGN.md §2 item 3)
      const float arg = __fmul_rn(s_abs, __fsub_rn(rd.w, 1.0f));
      s_e[tid] = (float)exp((double)arg);
    }
    __syncthreads();
    // ---- pooling weights per point (blending_network.py:76-80) ----
    if (tid < ppt) {
      const int r0 = tid * V;
      float emin = INFINITY;
      for (int v = 0; v < V; ++v) emin = fminf(emin, s_e[r0 + v]);
      float wsum = 0.f;
      for (int v = 0; v < V; ++v) wsum = __fadd_rn(wsum, __fmul_rn(__fsub_rn(s_e[r0 + v], emin), s_m[r0 + v]));
      const float den = __fadd_rn(wsum, 1e-8f);
      for (int v = 0; v < V; ++v) s_wv[r0 + v] = __fdiv_rn(__fmul_rn(__fsub_rn(s_e[r0 + v], emin), s_m[r0 + v]), den);
    }
    // ---- ray_dir_fc on the CUDA cores: thread (row, half): 4 -> 16 (all), then 19 outputs split 10 / 9 ----
    {
      float h16[16];
      const float d0 = RD[r], d1 = RD[BS + r], d2 = RD[2 * BS + r], d3 = RD[3 * BS + r];
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        const float* w = WF + F_DIR0_W + o * 4;
        h16[o] = bt_elu(fmaf(w[3], d3, fmaf(w[2], d2, fmaf(w[1], d1, fmaf(w[0], d0, WF[F_DIR0_B + o])))));
      }
      const int o0 = half ? 10 : 0, o1 = half ? 19 : 10;
      for (int o = o0; o < o1; ++o) {
        const float* w = WF + F_DIR1_W + o * 16;
        float a = WF[F_DIR1_B + o];
#pragma unroll
        for (int k = 0; k < 16; ++k) a = fmaf(w[k], h16[k], a);
        F[o * BS + r] += bt_elu(a);                       // x19 = feat19 + direction feature (in place)
      }
    }
    __syncthreads();
    // ---- A operand of base_fc.0: [mean19, var19, x19, 0 x 7]  (weighted mean / variance over the views, :15-19) ----
    // thread (row, half) builds 32 consecutive k of its own row in registers and stores them as 16-byte vectors
    // (conflict-free); the per-point statistics are recomputed by each of the point's V rows.
    {
      const int r0 = (r / V) * V;
      const bool live = r < rows;
      float vals[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) vals[k] = 0.f;
      if (live) {
        if (half == 0) {          // k 0..31 = mean 0..18, var 0..12
#pragma unroll
          for (int c = 0; c < 19; ++c) {
            const float* x = F + c * BS + r0;
            float mean = 0.f;
            for (int v = 0; v < V; ++v) mean = fmaf(x[v], s_wv[r0 + v], mean);
            vals[c] = mean;
            if (c < 13) {
              float var = 0.f;
              for (int v = 0; v < V; ++v) {
                const float d = x[v] - mean;
                var = fmaf(s_wv[r0 + v] * d, d, var);
              }
              vals[19 + c] = var;
            }
          }
        } else {                  // k 32..63 = var 13..18, x19, 0 x 7
#pragma unroll
          for (int c = 13; c < 19; ++c) {
            const float* x = F + c * BS + r0;
            float mean = 0.f;
            for (int v = 0; v < V; ++v) mean = fmaf(x[v], s_wv[r0 + v], mean);
            float var = 0.f;
            for (int v = 0; v < V; ++v) {
              const float d = x[v] - mean;
              var = fmaf(s_wv[r0 + v] * d, d, var);
            }
            vals[c - 13] = var;
          }
#pragma unroll
          for (int c = 0; c < 19; ++c) vals[6 + c] = F[c * BS + r];
        }
      }
      float lo16[16], hi16[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) { lo16[k] = vals[k]; hi16[k] = vals[16 + k]; }
      aop_store16(aop, r, half * 32, lo16);
      aop_store16(aop, r, half * 32 + 16, hi16);
    }
    publish();
    // ---- base_fc.0 : 57(64) -> 64, ELU ----
    if (warp == 0 && tc::elect_one()) bt_issue<64, 4>(tbase, aop_a, w_a + W_BASE0, &bars->mma_done, fast);
    wait_mma();
    {
      uint32_t acc[32];
      tc::tmem_ld32(tl + half * 32, acc);
      tc::tmem_wait_ld();
      float v[16];
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = bt_elu(__uint_as_float(acc[hb * 16 + j]) + WF[F_BASE0_B + half * 32 + hb * 16 + j]);
        aop_store16(aop, r, half * 32 + hb * 16, v);
      }
    }
    publish();
    // ---- base_fc.2 : 64 -> 32, ELU -> x (fp32) ; next A operand = x * pooling weight ----
    if (warp == 0 && tc::elect_one()) bt_issue<32, 4>(tbase, aop_a, w_a + W_BASE1, &bars->mma_done, fast);
    wait_mma();
    {
      uint32_t acc[16];
      tc::tmem_ld16(tl + half * 16, acc);
      tc::tmem_wait_ld();
      const float wv = s_wv[r];
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = bt_elu(__uint_as_float(acc[j]) + WF[F_BASE1_B + half * 16 + j]);
        XF[(half * 16 + j) * BS + r] = x;
        v[j] = x * wv;
      }
      aop_store16(aop, r, half * 16, v);
    }
    publish();
    // ---- vis_fc.0 : 32 -> 32, ELU ----
    if (warp == 0 && tc::elect_one()) bt_issue<32, 2>(tbase, aop_a, w_a + W_VIS0, &bars->mma_done, fast);
    wait_mma();
    {
      uint32_t acc[16];
      tc::tmem_ld16(tl + half * 16, acc);
      tc::tmem_wait_ld();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = bt_elu(__uint_as_float(acc[j]) + WF[F_VIS0_B + half * 16 + j]);
      aop_store16(aop, r, half * 16, v);
    }
    publish();
    // ---- vis_fc.2 : 32 -> 33, ELU ; x += x_res ; vis = sigmoid(.) * mask ----
    if (warp == 0 && tc::elect_one()) bt_issue<48, 2>(tbase, aop_a, w_a + W_VIS1, &bars->mma_done, fast);
    wait_mma();
    {
      uint32_t acc[16];
      tc::tmem_ld16(tl + half * 16, acc);
      if (half == 1) {
        uint32_t extra;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(extra) : "r"(tl + 32));
        tc::tmem_wait_ld();
        s_vis[r] = bt_sigmoid(bt_elu(__uint_as_float(extra) + WF[F_VIS1_B + 32])) * s_m[r];
      } else {
        tc::tmem_wait_ld();
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        XF[(half * 16 + j) * BS + r] += bt_elu(__uint_as_float(acc[j]) + WF[F_VIS1_B + half * 16 + j]);
    }
    tc::tc_fence_before();
    __syncthreads();
    {   // A operand of vis_fc2.0 = x * vis
      const float vs = s_vis[r];
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = XF[(half * 16 + j) * BS + r] * vs;
      aop_store16(aop, r, half * 16, v);
    }
    publish();
    // ---- vis_fc2.0 : 32 -> 32, ELU ; vis_fc2.2 : 32 -> 1 (CUDA cores), sigmoid * mask ----
    if (warp == 0 && tc::elect_one()) bt_issue<32, 2>(tbase, aop_a, w_a + W_V20, &bars->mma_done, fast);
    wait_mma();
    {
      uint32_t acc[16];
      tc::tmem_ld16(tl + half * 16, acc);
      tc::tmem_wait_ld();
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        p = fmaf(bt_elu(__uint_as_float(acc[j]) + WF[F_V20_B + half * 16 + j]), WF[F_V21_W + half * 16 + j], p);
      if (half == 1) s_part[r] = p;
      tc::tc_fence_before();
      __syncthreads();
      if (half == 0) s_vis[r] = bt_sigmoid(p + s_part[r] + WF[F_V21_B]) * s_m[r];      // vis2
    }
    __syncthreads();
    {   // A operand of rgb_fc.0 = [x (32), vis2, ray_diff (4), 0 x 11]  (K = 48)
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = XF[(half * 16 + j) * BS + r];
      aop_store16(aop, r, half * 16, v);
      if (half == 0) {
        float t[8] = {s_vis[r], RD[r], RD[BS + r], RD[2 * BS + r], RD[3 * BS + r], 0.f, 0.f, 0.f};
        aop_store8(aop, r, 32, t);
      } else {
        float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        aop_store8(aop, r, 40, t);
      }
    }
    publish();
    // ---- rgb_fc.0 : 37(48) -> 16, ELU ; rgb_fc.2/.4 : 16 -> 8 -> 1 on the CUDA cores ----
    if (warp == 0 && tc::elect_one()) bt_issue<16, 3>(tbase, aop_a, w_a + W_RGB0, &bars->mma_done, fast);
    wait_mma();
    {
      float* A16 = F;                                   // [16][BS], feat buffer is free now
      uint32_t acc[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
                     "=r"(acc[7])
                   : "r"(tl + half * 8));
      tc::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) A16[(half * 8 + j) * BS + r] = bt_elu(__uint_as_float(acc[j]) + WF[F_RGB0_B + half * 8 + j]);
      tc::tc_fence_before();
      __syncthreads();
      if (half == 0) {
        float a16[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a16[k] = A16[k * BS + r];
        float lg = WF[F_RGB2_B];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const float* w = WF + F_RGB1_W + o * 16;
          float a = WF[F_RGB1_B + o];
#pragma unroll
          for (int k = 0; k < 16; ++k) a = fmaf(w[k], a16[k], a);
          lg = fmaf(bt_elu(a), WF[F_RGB2_W + o], lg);
        }
        s_logit[r] = lg;
      }
    }
    __syncthreads();
    // ---- masked softmax over views, blend the source RGB (:109-115) ----
    if (tid < ppt) {
      const int64_t i = tile * ppt + tid;
      if (i < n_total) {
        const int r0 = tid * V;
        float mx = -INFINITY;
        unsigned vbits = 0;
        for (int v = 0; v < V; ++v) {
          const bool ok = s_m[r0 + v] > 0.f;
          if (ok) vbits |= 1u << v;
          mx = fmaxf(mx, ok ? s_logit[r0 + v] : -1e9f);
        }
        float ssum = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
        for (int v = 0; v < V; ++v) {
          const float lg = (s_m[r0 + v] > 0.f) ? s_logit[r0 + v] : -1e9f;
          const float e = expf(lg - mx);
          ssum += e;
          cr = fmaf(e, RGB[r0 + v], cr);
          cg = fmaf(e, RGB[BS + r0 + v], cg);
          cb = fmaf(e, RGB[2 * BS + r0 + v], cb);
        }
        const int64_t id = list ? (int64_t)list[i] : i;
        const float inv = 1.0f / ssum;
        rgb_out[id * 3] = cr * inv;
        rgb_out[id * 3 + 1] = cg * inv;
        rgb_out[id * 3 + 2] = cb * inv;
        if (views_out) views_out[id] = (uint8_t)vbits;
      }
    }
    __syncthreads();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<64>(tbase);
}

// =============================================================================================
// k_blend_tm: the same network with THREE tiles in flight per CTA and no block-wide barrier (round 2).
//
// k_blend_tc above is block-synchronous: every layer is "publish (fence + __syncthreads) -> one thread issues -> all
// wait for the MMAs -> epilogue", ~16 block barriers and 6 exposed MMA round trips per tile with only two tiles per SM
// (ncu r02: tensor pipe 10.6 %, issue slots 47.7 %, 1.6 barrier + 1.8 scoreboard stalls per issue).  Here
//   * activations never touch shared memory: the epilogue writes the next A operand (fp16 hi | lo) into TMEM with
//     tcgen05.st and the MMAs take A from TMEM (TS form), like the SDF kernel; the 32-channel residual stream x lives
//     in registers (16 columns per thread);
//   * a CTA runs three independent tile groups of 8 warps (warp (q, half): TMEM lane quarter q, column half) + one
//     issuer warp; each group owns 128 TMEM columns (A hi 32 | A lo 32 | D 64) and two mbarriers (a_ready: 8 warp
//     arrivals, d_full: tcgen05.commit), so one group's MMA round trip is hidden behind the other groups' epilogues;
//   * the only thread synchronisation left is between the two warps that share a lane quarter (named barrier of 64
//     threads) for the few values that cross the column halves; the V rows of a point are adjacent lanes of one warp,
//     so the cross-view operations of the output stage (masked softmax, RGB blend) are warp shuffles.
// V in {2, 4} (the rows of a point must not straddle a warp); other view counts use k_blend_tc.  Same arithmetic as
// k_blend_tc (the same formulas in the same order per value), hence the same parity envelope.
// =============================================================================================
#define TM_GROUPS 3
#define TM_EPI_WARPS 8
#define TM_THREADS ((TM_GROUPS * TM_EPI_WARPS + 1) * 32)
#define TMF_THREADS (TM_THREADS + TM_GROUPS * 32)   // fused edition: + one projection-gather warp per group
#define TM_REC_CH 24                                // staged record per row: 20 (FEAT_REC) + ray_diff 4
#define TM_AHI(g) ((uint32_t)(g) * 128u)
#define TM_ALO(g) ((uint32_t)(g) * 128u + 32u)
#define TM_D(g) ((uint32_t)(g) * 128u + 64u)

struct TmGroup {
  float X[19][BS];        // x19 = feat19 + direction feature
  float A16[16][BS];      // rgb_fc.0 outputs crossing the halves
  float RGB[3][BS];
  float RD[4][BS];
  float m[128], e[128], vis[128], part[128];
};
struct TmBars {
  uint64_t a_ready[TM_GROUPS];
  uint64_t d_full[TM_GROUPS];
  uint64_t rec_full[TM_GROUPS * 2];    // fused edition: record stage (group, buffer) written by the gather warp
  uint64_t rec_empty[TM_GROUPS * 2];   //                ... and read by all 8 epilogue warps of the group
  uint32_t tmem_base;
};
#define SM_TM_WF W_TC_BYTES
#define SM_TM_GRP ((SM_TM_WF + F_TOTAL * 4 + 127) / 128 * 128)
#define SM_TM_BAR (SM_TM_GRP + TM_GROUPS * (int)sizeof(TmGroup))
#define SM_TM_TOTAL (SM_TM_BAR + 256)
#define SM_TM_REC SM_TM_TOTAL        // fused edition: [group][2][TM_REC_CH][BS] fp32 record stages
#define SM_TMF_TOTAL (SM_TM_REC + TM_GROUPS * 2 * TM_REC_CH * BS * 4)
static_assert(sizeof(TmBars) <= 256, "barrier block");
static_assert(SM_TMF_TOTAL <= 227 * 1024, "shared memory of the fused colour kernel");

// 16 values (consecutive k, k0 % 16 == 0) -> hi / lo words at TMEM columns k0/2 .. k0/2+7 of the A regions
__device__ __forceinline__ void tm_store16(uint32_t t_hi, uint32_t t_lo, int k0, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) tc::split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  tc::tmem_st8(t_hi + (k0 >> 1), hi);
  tc::tmem_st8(t_lo + (k0 >> 1), lo);
}
__device__ __forceinline__ void tm_store8(uint32_t t_hi, uint32_t t_lo, int k0, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) tc::split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  tc::tmem_st4(t_hi + (k0 >> 1), hi);
  tc::tmem_st4(t_lo + (k0 >> 1), lo);
}
// one layer: D[128 x N] = A[128 x K] (TMEM, hi | lo) * W^T (smem), issued by one thread
template <int N, int KSTEPS>
__device__ __forceinline__ void tm_issue(uint32_t tD, uint32_t a_hi, uint32_t a_lo, uint32_t w_addr, uint64_t* bar, bool fast) {
  const uint32_t idesc = tc::idesc_f16(128, N, 0);
  const uint64_t db = tc::smem_desc_kmajor(0, N * 16, 128);
  const uint32_t bh = (uint32_t)(db >> 32);
  const uint32_t b0 = (uint32_t)db | (w_addr >> 4);
  constexpr uint32_t B_LO = (N * KSTEPS * 16 * 2) >> 4, B_KS = (N * 32) >> 4;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    if (ks == 0) tc::mma_ts_w<false>(tD, a_hi, b0, bh, idesc);
    else tc::mma_ts_w<true>(tD, a_hi + ks * 8, b0 + ks * B_KS, bh, idesc);
    if (!fast) tc::mma_ts_w<true>(tD, a_lo + ks * 8, b0 + ks * B_KS, bh, idesc);
    if (!fast) tc::mma_ts_w<true>(tD, a_hi + ks * 8, b0 + B_LO + ks * B_KS, bh, idesc);
  }
  tc::mma_commit(bar);
}

// FUSED: the records are not read from HBM but produced in the kernel: one extra warp per tile group projects the
// group's next tile into the source views and gathers RGB + the feature pyramid (lookup_row, the body of
// k_lookup_feature) into a double-buffered shared-memory stage while the group's 8 warps run the network on the
// current tile.  Removes the 96 B / (point, view) HBM round trip and the separate gather kernel.
template <int V, bool FUSED>
__global__ void __launch_bounds__(FUSED ? TMF_THREADS : TM_THREADS, 1)
k_blend_tm(const uint8_t* __restrict__ wtc, const float* __restrict__ wf32, float s_abs, const float* __restrict__ feat,
           const float* __restrict__ rdiff, const uint8_t* __restrict__ mask, int packed19,
           const int32_t* __restrict__ list, const int32_t* __restrict__ count, int64_t n, float* __restrict__ rgb_out,
           uint8_t* __restrict__ views_out, int fast_i, const DevScene sc, const PointSource src) {
  const bool fast = fast_i != 0;
  constexpr int NTHREADS = FUSED ? TMF_THREADS : TM_THREADS;
  extern __shared__ __align__(1024) uint8_t smem[];
  TmBars* bars = reinterpret_cast<TmBars*>(smem + SM_TM_BAR);
  const float* WF = reinterpret_cast<const float*>(smem + SM_TM_WF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < W_TC_BYTES / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(wtc)[i];
  for (int i = tid; i < F_TOTAL; i += NTHREADS) reinterpret_cast<float*>(smem + SM_TM_WF)[i] = wf32[i];
  if (warp == TM_GROUPS * TM_EPI_WARPS) tc::tmem_alloc<512>(&bars->tmem_base);
  if (tid == 0) {
    for (int g = 0; g < TM_GROUPS; ++g) {
      tc::mbar_init(&bars->a_ready[g], TM_EPI_WARPS);
      tc::mbar_init(&bars->d_full[g], 1);
    }
    for (int i = 0; i < TM_GROUPS * 2; ++i) {
      tc::mbar_init(&bars->rec_full[i], 1);
      tc::mbar_init(&bars->rec_empty[i], TM_EPI_WARPS);
    }
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tbase = bars->tmem_base;

  int64_t n_total = n;
  if (count) {
    const int64_t c = *count;
    n_total = c < n_total ? c : n_total;
  }
  constexpr int ppt = BT_ROWS / V;
  const int64_t n_tiles = (n_total + ppt - 1) / ppt;
  const int64_t stride = (int64_t)gridDim.x * TM_GROUPS;

  if (warp < TM_GROUPS * TM_EPI_WARPS) {
    // =============================== epilogue warps ===============================
    const int g = warp / TM_EPI_WARPS, lw = warp % TM_EPI_WARPS;
    const int q = lw & 3, half = lw >> 2;
    const int r = q * 32 + lane;
    TmGroup& G = *reinterpret_cast<TmGroup*>(smem + SM_TM_GRP + g * sizeof(TmGroup));
    const uint32_t tl = tbase + ((uint32_t)(q * 32) << 16);
    const uint32_t t_hi = tl + TM_AHI(g), t_lo = tl + TM_ALO(g), t_d = tl + TM_D(g);
    const int pair_bar = 1 + g * 4 + q;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    auto signal = [&]() {           // my part of the next A operand is in TMEM
      tc::tmem_wait_st();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->a_ready[g]);
    };
    uint32_t ph = 0;
    auto wait_d = [&]() {
      tc::mbar_wait(&bars->d_full[g], ph & 1);
      ph++;
      tc::tc_fence_after();
    };
    uint32_t kt = 0;               // tiles this group has taken (record stage kt & 1, phase kt >> 1)
    for (int64_t tile = (int64_t)blockIdx.x * TM_GROUPS + g; tile < n_tiles; tile += stride, ++kt) {
      // ---- record of my row: row r <-> (point tile*ppt + r / V, view r % V) ----
      const int64_t rec_i = tile * BT_ROWS + r;
      const bool ok = rec_i < n_total * V;
      float rec[FEAT_REC];
      float4 rd = make_float4(0.f, 0.f, 0.f, 1.f);
#pragma unroll
      for (int c = 0; c < FEAT_REC; ++c) rec[c] = 0.f;
      const int stage = g * 2 + (int)(kt & 1);
      const float(*R)[BS] = reinterpret_cast<const float(*)[BS]>(smem + SM_TM_REC + stage * (TM_REC_CH * BS * 4));
      if (FUSED) {       // the record comes from the group's gather warp through shared memory
        tc::mbar_wait(&bars->rec_full[stage], (kt >> 1) & 1);
        rd = make_float4(R[20][r], R[21][r], R[22][r], R[23][r]);
      } else if (ok) {
        rd = reinterpret_cast<const float4*>(rdiff)[rec_i];
        if (packed19) {
          const float* f = feat + rec_i * 19;
#pragma unroll
          for (int c = 0; c < 19; ++c) rec[c] = f[c];
          rec[19] = mask[rec_i] ? 1.0f : 0.0f;
        } else {
          const float4* f = reinterpret_cast<const float4*>(feat + rec_i * FEAT_REC);
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            const float4 t = f[c];
            rec[4 * c] = t.x; rec[4 * c + 1] = t.y; rec[4 * c + 2] = t.z; rec[4 * c + 3] = t.w;
          }
        }
      }
      // ---- ray_dir_fc 4 -> 16 (both halves), 16 -> 19 split 10 / 9 ; x19 -> smem ----
      {
        float h16[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          const float* w = WF + F_DIR0_W + o * 4;
          h16[o] = bt_elu(fmaf(w[3], rd.w, fmaf(w[2], rd.z, fmaf(w[1], rd.y, fmaf(w[0], rd.x, WF[F_DIR0_B + o])))));
        }
        const int o0 = half ? 10 : 0, o1 = half ? 19 : 10;
        for (int o = o0; o < o1; ++o) {
          const float* w = WF + F_DIR1_W + o * 16;
          float a = WF[F_DIR1_B + o];
#pragma unroll
          for (int k = 0; k < 16; ++k) a = fmaf(w[k], h16[k], a);
          G.X[o][r] = (FUSED ? R[o][r] : rec[o]) + bt_elu(a);
        }
      }
      if (half == 0) {
        if (FUSED) {
          G.RGB[0][r] = R[0][r]; G.RGB[1][r] = R[1][r]; G.RGB[2][r] = R[2][r];
          G.m[r] = R[19][r];
        } else {
          G.RGB[0][r] = rec[0]; G.RGB[1][r] = rec[1]; G.RGB[2][r] = rec[2];
          G.m[r] = rec[19];
        }
        G.RD[0][r] = rd.x; G.RD[1][r] = rd.y; G.RD[2][r] = rd.z; G.RD[3][r] = rd.w;
        // pooling exponential, correctly rounded (see blend.cu / DESIGN.md §2)
        const float arg = __fmul_rn(s_abs, __fsub_rn(rd.w, 1.0f));
        G.e[r] = (float)exp((double)arg);
      }
      if (FUSED) {       // my reads of the record stage are done: hand it back to the gather warp
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->rec_empty[stage]);
      }
      pair_sync();
      // ---- pooling weights of my point (blending_network.py:76-80) ----
      const int r0 = (r / V) * V;
      float wvs[V];
      {
        float emin = INFINITY;
#pragma unroll
        for (int v = 0; v < V; ++v) emin = fminf(emin, G.e[r0 + v]);
        float wsum = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) wsum = __fadd_rn(wsum, __fmul_rn(__fsub_rn(G.e[r0 + v], emin), G.m[r0 + v]));
        const float den = __fadd_rn(wsum, 1e-8f);
#pragma unroll
        for (int v = 0; v < V; ++v) wvs[v] = __fdiv_rn(__fmul_rn(__fsub_rn(G.e[r0 + v], emin), G.m[r0 + v]), den);
      }
      const float wv = wvs[r - r0];
      const float mrow = G.m[r];
      // ---- A operand of base_fc.0: [mean19, var19, x19, 0 x 7] ----
      {
        float vals[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) vals[k] = 0.f;
        if (half == 0) {
#pragma unroll
          for (int c = 0; c < 19; ++c) {
            float mean = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) mean = fmaf(G.X[c][r0 + v], wvs[v], mean);
            vals[c] = mean;
            if (c < 13) {
              float var = 0.f;
#pragma unroll
              for (int v = 0; v < V; ++v) {
                const float d = G.X[c][r0 + v] - mean;
                var = fmaf(wvs[v] * d, d, var);
              }
              vals[19 + c] = var;
            }
          }
        } else {
#pragma unroll
          for (int c = 13; c < 19; ++c) {
            float mean = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) mean = fmaf(G.X[c][r0 + v], wvs[v], mean);
            float var = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float d = G.X[c][r0 + v] - mean;
              var = fmaf(wvs[v] * d, d, var);
            }
            vals[c - 13] = var;
          }
#pragma unroll
          for (int c = 0; c < 19; ++c) vals[6 + c] = G.X[c][r];
        }
        float a[16], b[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { a[k] = vals[k]; b[k] = vals[16 + k]; }
        tm_store16(t_hi, t_lo, half * 32, a);
        tm_store16(t_hi, t_lo, half * 32 + 16, b);
      }
      signal();
      // ---- base_fc.0 : 57(64) -> 64, ELU ----
      wait_d();
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        uint32_t acc[16];
        tc::tmem_ld16(t_d + half * 32 + hb * 16, acc);
        tc::tmem_wait_ld();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = bt_elu(__uint_as_float(acc[j]) + WF[F_BASE0_B + half * 32 + hb * 16 + j]);
        tm_store16(t_hi, t_lo, half * 32 + hb * 16, v);
      }
      signal();
      // ---- base_fc.2 : 64 -> 32, ELU -> x (registers) ; next A = x * pooling weight ----
      float x[16];
      wait_d();
      {
        uint32_t acc[16];
        tc::tmem_ld16(t_d + half * 16, acc);
        tc::tmem_wait_ld();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          x[j] = bt_elu(__uint_as_float(acc[j]) + WF[F_BASE1_B + half * 16 + j]);
          v[j] = x[j] * wv;
        }
        tm_store16(t_hi, t_lo, half * 16, v);
      }
      signal();
      // ---- vis_fc.0 : 32 -> 32, ELU ----
      wait_d();
      {
        uint32_t acc[16];
        tc::tmem_ld16(t_d + half * 16, acc);
        tc::tmem_wait_ld();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = bt_elu(__uint_as_float(acc[j]) + WF[F_VIS0_B + half * 16 + j]);
        tm_store16(t_hi, t_lo, half * 16, v);
      }
      signal();
      // ---- vis_fc.2 : 32 -> 33, ELU ; x += x_res ; vis = sigmoid(.) * mask ; next A = x * vis ----
      wait_d();
      {
        uint32_t acc[16];
        tc::tmem_ld16(t_d + half * 16, acc);
        if (half == 1) {
          uint32_t extra;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(extra) : "r"(t_d + 32));
          tc::tmem_wait_ld();
          G.vis[r] = bt_sigmoid(bt_elu(__uint_as_float(extra) + WF[F_VIS1_B + 32])) * mrow;
        } else {
          tc::tmem_wait_ld();
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] += bt_elu(__uint_as_float(acc[j]) + WF[F_VIS1_B + half * 16 + j]);
        pair_sync();
        const float vs = G.vis[r];
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = x[j] * vs;
        tm_store16(t_hi, t_lo, half * 16, v);
      }
      signal();
      // ---- vis_fc2.0 : 32 -> 32, ELU ; vis_fc2.2 : 32 -> 1, sigmoid * mask ; next A = [x, vis2, ray_diff, 0] ----
      wait_d();
      {
        uint32_t acc[16];
        tc::tmem_ld16(t_d + half * 16, acc);
        tc::tmem_wait_ld();
        float p = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          p = fmaf(bt_elu(__uint_as_float(acc[j]) + WF[F_V20_B + half * 16 + j]), WF[F_V21_W + half * 16 + j], p);
        if (half == 1) G.part[r] = p;
        pair_sync();
        tm_store16(t_hi, t_lo, half * 16, x);
        if (half == 0) {
          const float vis2 = bt_sigmoid(p + G.part[r] + WF[F_V21_B]) * mrow;
          const float t[8] = {vis2, G.RD[0][r], G.RD[1][r], G.RD[2][r], G.RD[3][r], 0.f, 0.f, 0.f};
          tm_store8(t_hi, t_lo, 32, t);
        } else {
          const float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          tm_store8(t_hi, t_lo, 40, t);
        }
      }
      signal();
      // ---- rgb_fc.0 : 37(48) -> 16, ELU ; rgb_fc.2/.4 : 16 -> 8 -> 1 ; masked softmax over the views ; blend ----
      wait_d();
      {
        uint32_t acc[8];
        tc::tmem_ld8(t_d + half * 8, acc);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) G.A16[half * 8 + j][r] = bt_elu(__uint_as_float(acc[j]) + WF[F_RGB0_B + half * 8 + j]);
        // the accumulator has been read: the group's next tile may start while the output stage runs
        tc::tc_fence_before();
        pair_sync();
        if (half == 0) {
          float a16[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) a16[k] = G.A16[k][r];
          float lg = WF[F_RGB2_B];
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const float* w = WF + F_RGB1_W + o * 16;
            float a = WF[F_RGB1_B + o];
#pragma unroll
            for (int k = 0; k < 16; ++k) a = fmaf(w[k], a16[k], a);
            lg = fmaf(bt_elu(a), WF[F_RGB2_W + o], lg);
          }
          // masked softmax over the V adjacent lanes of my point, in view order (blending_network.py:109-115)
          const bool live = mrow > 0.f;
          const float lgm = live ? lg : -1e9f;
          float mx = lgm;
#pragma unroll
          for (int o = 1; o < V; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          const float ex = expf(lgm - mx);
          const int base_lane = lane - (lane % V);
          float ssum = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
          unsigned vbits = 0;
#pragma unroll
          for (int v = 0; v < V; ++v) {       // sequential over the views like the scalar loop of k_blend_tc
            const float ev = __shfl_sync(0xffffffffu, ex, base_lane + v);
            const unsigned lv = __shfl_sync(0xffffffffu, live ? 1u : 0u, base_lane + v);
            vbits |= lv << v;
            ssum += ev;
            cr = fmaf(ev, G.RGB[0][r0 + v], cr);
            cg = fmaf(ev, G.RGB[1][r0 + v], cg);
            cb = fmaf(ev, G.RGB[2][r0 + v], cb);
          }
          if (r == r0) {
            const int64_t i = tile * ppt + r / V;
            if (i < n_total) {
              const int64_t id = list ? (int64_t)list[i] : i;
              const float inv = 1.0f / ssum;
              rgb_out[id * 3] = cr * inv;
              rgb_out[id * 3 + 1] = cg * inv;
              rgb_out[id * 3 + 2] = cb * inv;
              if (views_out) views_out[id] = (uint8_t)vbits;
            }
          }
        }
      }
    }
  } else if (FUSED && warp > TM_GROUPS * TM_EPI_WARPS) {
    // =============================== projection-gather warps (one per group) ===============================
    const int g = warp - TM_GROUPS * TM_EPI_WARPS - 1;
    uint32_t kt = 0;
    for (int64_t tile = (int64_t)blockIdx.x * TM_GROUPS + g; tile < n_tiles; tile += stride, ++kt) {
      const int stage = g * 2 + (int)(kt & 1);
      float(*R)[BS] = reinterpret_cast<float(*)[BS]>(smem + SM_TM_REC + stage * (TM_REC_CH * BS * 4));
      tc::mbar_wait(&bars->rec_empty[stage], ((kt >> 1) & 1) ^ 1);
      // phase 1: the sample positions of my 4 rows (the dependent loads list -> mid_z / ray of the rows overlap),
      // parked in the ray_diff slots of the stage
#pragma unroll
      for (int j = 0; j < BT_ROWS / 32; ++j) {
        const int row = j * 32 + lane;
        const int64_t rec_i = tile * BT_ROWS + row;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (rec_i < n_total * V) {
          const int64_t i = rec_i / V;
          if (src.mode == 0) {
            px = src.pts[i * 3]; py = src.pts[i * 3 + 1]; pz = src.pts[i * 3 + 2];
          } else {
            const int64_t id = src.list ? (int64_t)src.list[i] : i;
            const int64_t ray = id / src.S;
            const float t = src.mid_z[id];
            px = ray_at(src.rays_o[ray * 3], src.rays_d[ray * 3], t);
            py = ray_at(src.rays_o[ray * 3 + 1], src.rays_d[ray * 3 + 1], t);
            pz = ray_at(src.rays_o[ray * 3 + 2], src.rays_d[ray * 3 + 2], t);
          }
        }
        R[20][row] = px; R[21][row] = py; R[22][row] = pz;
      }
      // phase 2: projection + gather, one row at a time (each lane only re-reads what it wrote)
#pragma unroll 1
      for (int j = 0; j < BT_ROWS / 32; ++j) {
        const int row = j * 32 + lane;
        const int64_t rec_i = tile * BT_ROWS + row;
        float rec[FEAT_REC];
        float4 rd = make_float4(0.f, 0.f, 0.f, 1.f);
        if (rec_i < n_total * V) {
          lookup_row(sc, R[20][row], R[21][row], R[22][row], (int)(rec_i % V), rec, rd);
        } else {
#pragma unroll
          for (int c = 0; c < FEAT_REC; ++c) rec[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < FEAT_REC; ++c) R[c][row] = rec[c];
        R[20][row] = rd.x; R[21][row] = rd.y; R[22][row] = rd.z; R[23][row] = rd.w;
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->rec_full[stage]);
    }
  } else if (warp == TM_GROUPS * TM_EPI_WARPS) {
    // =============================== the MMA issuer ===============================
    if (tc::elect_one()) {
      const uint32_t w_a = tc::smem_u32(smem);
      // round robin over the groups: whichever has its next A operand ready gets its layer issued (a blocking wait on
      // one group would hold back the MMAs of the other two)
      uint32_t ph[TM_GROUPS];
      int layer[TM_GROUPS];
      int64_t tile_of[TM_GROUPS];
      int live = 0;
#pragma unroll
      for (int g = 0; g < TM_GROUPS; ++g) {
        ph[g] = 0;
        layer[g] = 0;
        tile_of[g] = (int64_t)blockIdx.x * TM_GROUPS + g;
        live += tile_of[g] < n_tiles ? 1 : 0;
      }
      while (live > 0) {
#pragma unroll
        for (int g = 0; g < TM_GROUPS; ++g) {
          if (tile_of[g] >= n_tiles) continue;
          if (!tc::mbar_try_wait(&bars->a_ready[g], ph[g] & 1)) continue;
          ph[g]++;
          tc::tc_fence_after();
          const uint32_t tD = tbase + TM_D(g), ah = tbase + TM_AHI(g), al = tbase + TM_ALO(g);
          switch (layer[g]) {
            case 0: tm_issue<64, 4>(tD, ah, al, w_a + W_BASE0, &bars->d_full[g], fast); break;
            case 1: tm_issue<32, 4>(tD, ah, al, w_a + W_BASE1, &bars->d_full[g], fast); break;
            case 2: tm_issue<32, 2>(tD, ah, al, w_a + W_VIS0, &bars->d_full[g], fast); break;
            case 3: tm_issue<48, 2>(tD, ah, al, w_a + W_VIS1, &bars->d_full[g], fast); break;
            case 4: tm_issue<32, 2>(tD, ah, al, w_a + W_V20, &bars->d_full[g], fast); break;
            default: tm_issue<16, 3>(tD, ah, al, w_a + W_RGB0, &bars->d_full[g], fast); break;
          }
          if (++layer[g] == 6) {
            layer[g] = 0;
            tile_of[g] += stride;
            if (tile_of[g] >= n_tiles) --live;
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == TM_GROUPS * TM_EPI_WARPS) tc::tmem_dealloc<512>(tbase);
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static inline uint16_t bt_f2h(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
static inline float bt_h2f(uint16_t b) {
  __half h;
  memcpy(&h, &b, 2);
  return __half2float(h);
}

// W (out, in) row-major fp32 -> [hi | lo] canonical K-major operand with N rows (>= out), K columns (>= in)
static void bt_put_matrix(std::vector<uint8_t>& blob, int off, const float* W, int out, int in, int N, int K) {
  uint16_t* hi = reinterpret_cast<uint16_t*>(blob.data() + off);
  uint16_t* lo = hi + (size_t)N * K;
  for (int n = 0; n < out; ++n)
    for (int k = 0; k < in; ++k) {
      const float v = W[(size_t)n * in + k];
      const uint16_t h = bt_f2h(v);
      const size_t idx = (size_t)(k >> 3) * N * 8 + (size_t)n * 8 + (k & 7);
      hi[idx] = h;
      lo[idx] = bt_f2h(v - bt_h2f(h));
    }
}

int surf_build_blend_tc_weights(const surf_net_inputs* in, surf_net* net, cudaStream_t st,
                                int (*dev_alloc)(surf_net*, void**, size_t)) {
  // h_blend_w order: ray_dir_fc.0,.2 base_fc.0,.2 vis_fc.0,.2 vis_fc2.0,.2 rgb_fc.0,.2,.4
  std::vector<uint8_t> tcw(W_TC_BYTES, 0);
  bt_put_matrix(tcw, W_BASE0, in->h_blend_w[2], 64, 57, 64, 64);
  bt_put_matrix(tcw, W_BASE1, in->h_blend_w[3], 32, 64, 32, 64);
  bt_put_matrix(tcw, W_VIS0, in->h_blend_w[4], 32, 32, 32, 32);
  bt_put_matrix(tcw, W_VIS1, in->h_blend_w[5], 33, 32, 48, 32);
  bt_put_matrix(tcw, W_V20, in->h_blend_w[6], 32, 32, 32, 32);
  bt_put_matrix(tcw, W_RGB0, in->h_blend_w[8], 16, 37, 16, 48);
  std::vector<float> wf(F_TOTAL, 0.f);
  auto cp = [&](int off, const float* src, int n) { for (int i = 0; i < n; ++i) wf[off + i] = src[i]; };
  cp(F_DIR0_W, in->h_blend_w[0], 64);  cp(F_DIR0_B, in->h_blend_b[0], 16);
  cp(F_DIR1_W, in->h_blend_w[1], 304); cp(F_DIR1_B, in->h_blend_b[1], 19);
  cp(F_BASE0_B, in->h_blend_b[2], 64); cp(F_BASE1_B, in->h_blend_b[3], 32);
  cp(F_VIS0_B, in->h_blend_b[4], 32);  cp(F_VIS1_B, in->h_blend_b[5], 33);
  cp(F_V20_B, in->h_blend_b[6], 32);
  cp(F_V21_W, in->h_blend_w[7], 32);   cp(F_V21_B, in->h_blend_b[7], 1);
  cp(F_RGB0_B, in->h_blend_b[8], 16);
  cp(F_RGB1_W, in->h_blend_w[9], 128); cp(F_RGB1_B, in->h_blend_b[9], 8);
  cp(F_RGB2_W, in->h_blend_w[10], 8);  cp(F_RGB2_B, in->h_blend_b[10], 1);
  void* p = nullptr;
  int rc = dev_alloc(net, &p, tcw.size());
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, tcw.data(), tcw.size(), cudaMemcpyHostToDevice, st));
  net->blend_tc_w = (const uint8_t*)p;
  rc = dev_alloc(net, &p, wf.size() * sizeof(float));
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  net->blend_tc_f = (const float*)p;
  SURF_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int launch_blend_tc(const surf_net* n, const float* d_feat, const float* d_raydiff, const uint8_t* d_mask, int V,
                    bool packed19, const int32_t* list, const int32_t* count, int64_t n_pts, float* d_rgb,
                    uint8_t* d_views, bool fast, cudaStream_t st) {
  if (n_pts <= 0) return 0;
  {
    const void* fns[4] = {(const void*)k_blend_tc<0>, (const void*)k_blend_tc<2>, (const void*)k_blend_tc<3>,
                          (const void*)k_blend_tc<4>};
    for (int i = 0; i < 4; ++i) {
      const int rc = surf_ensure_dyn_smem(fns[i], SB_TOTAL);
      if (rc) return rc;
    }
  }
  const int ppt = BT_ROWS / V;
  const int64_t tiles = (n_pts + ppt - 1) / ppt;
  if (V == 2 || V == 4) {        // three tiles in flight per CTA, activations in TMEM, no block barriers
    int rc = surf_ensure_dyn_smem((const void*)k_blend_tm<2, false>, SM_TM_TOTAL);
    if (rc) return rc;
    rc = surf_ensure_dyn_smem((const void*)k_blend_tm<4, false>, SM_TM_TOTAL);
    if (rc) return rc;
    const int64_t want = (tiles + TM_GROUPS - 1) / TM_GROUPS;
    const int grid = (int)(want < n->n_sm ? want : n->n_sm);
    DevScene no_scene;
    PointSource no_src;
    memset(&no_scene, 0, sizeof(no_scene));
    memset(&no_src, 0, sizeof(no_src));
    surf_time_begin(3, st);
    if (V == 2)
      k_blend_tm<2, false><<<grid, TM_THREADS, SM_TM_TOTAL, st>>>(n->blend_tc_w, n->blend_tc_f, n->dev.blend_s, d_feat,
                                                                 d_raydiff, d_mask, packed19 ? 1 : 0, list, count, n_pts,
                                                                 d_rgb, d_views, fast ? 1 : 0, no_scene, no_src);
    else
      k_blend_tm<4, false><<<grid, TM_THREADS, SM_TM_TOTAL, st>>>(n->blend_tc_w, n->blend_tc_f, n->dev.blend_s, d_feat,
                                                                 d_raydiff, d_mask, packed19 ? 1 : 0, list, count, n_pts,
                                                                 d_rgb, d_views, fast ? 1 : 0, no_scene, no_src);
    surf_time_end(3, st);
    SURF_LAUNCH_CHECK();
    return 0;
  }
  const int64_t cap = (int64_t)n->n_sm * 2;
  const int grid = (int)(tiles < cap ? tiles : cap);
  surf_time_begin(3, st);
#define BT_LAUNCH(VT)                                                                                                  \
  k_blend_tc<VT><<<grid, BT_THREADS, SB_TOTAL, st>>>(n->blend_tc_w, n->blend_tc_f, n->dev.blend_s, d_feat, d_raydiff,   \
                                                     d_mask, V, packed19 ? 1 : 0, list, count, n_pts, d_rgb, d_views,  \
                                                     fast ? 1 : 0)
  switch (V) {
    case 2: BT_LAUNCH(2); break;
    case 3: BT_LAUNCH(3); break;
    case 4: BT_LAUNCH(4); break;
    default: BT_LAUNCH(0); break;
  }
#undef BT_LAUNCH
  surf_time_end(3, st);
  SURF_LAUNCH_CHECK();
  return 0;
}

// projection gather + blending network in one kernel (render path, V in {2, 4}, tensor-core modes)
bool color_fused_supported(const surf_scene* s, const surf_net* n, int mode) {
  return mode != SURF_MLP_FFMA && n->blend_tc_w != nullptr && (s->dev.V == 2 || s->dev.V == 4);
}

int launch_color_fused(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_rgb, uint8_t* d_views,
                       bool fast, cudaStream_t st) {
  SURF_CHECK_ARG(s->dev.img0, "scene has no images / feature maps");
  if (src.n <= 0) return 0;
  const int V = s->dev.V;
  int rc = surf_ensure_dyn_smem((const void*)k_blend_tm<2, true>, SM_TMF_TOTAL);
  if (rc) return rc;
  rc = surf_ensure_dyn_smem((const void*)k_blend_tm<4, true>, SM_TMF_TOTAL);
  if (rc) return rc;
  const int ppt = BT_ROWS / V;
  const int64_t tiles = (src.n + ppt - 1) / ppt;
  const int64_t want = (tiles + TM_GROUPS - 1) / TM_GROUPS;
  const int grid = (int)(want < n->n_sm ? want : n->n_sm);
  surf_time_begin(3, st);
  if (V == 2)
    k_blend_tm<2, true><<<grid, TMF_THREADS, SM_TMF_TOTAL, st>>>(n->blend_tc_w, n->blend_tc_f, n->dev.blend_s, nullptr,
                                                                nullptr, nullptr, 0, src.list, src.count, src.n, d_rgb,
                                                                d_views, fast ? 1 : 0, s->dev, src);
  else
    k_blend_tm<4, true><<<grid, TMF_THREADS, SM_TMF_TOTAL, st>>>(n->blend_tc_w, n->blend_tc_f, n->dev.blend_s, nullptr,
                                                                nullptr, nullptr, 0, src.list, src.count, src.n, d_rgb,
                                                                d_views, fast ? 1 : 0, s->dev, src);
  surf_time_end(3, st);
  SURF_LAUNCH_CHECK();
  return 0;
}
