// K2a + K3 (fp32 FFMA edition): sparse trilinear gather + SDF MLP forward + analytic input gradient.
//   reference: SDFNetworkSparse.forward / .sdf / .gradient (sdf_network.py:95-141),
//              lookup_sparse_volume (projector.py:217-390), Embedder (embedder.py:6-51).
//
// One persistent CTA per SM works on tiles of 128 points.  Activations stay in shared memory for the
// whole forward + reverse pass (k-major tile A[k][pt]); the 7 + 6 weight matrices are streamed from L2
// through a 3-slot cp.async ring as 32-row chunks in one fixed order ("weight stream"), so the kernel
// is a single uniform software pipeline.  Each thread owns an 8 (points) x 8 (outputs) register tile
// (8 x 10 in the reverse pass: 128 hidden + 28 feature-gradient columns, the latter accumulated in
// registers across layers).  softplus'(z) of every hidden layer is parked in a per-CTA, thread-private
// scratch (L2 resident) between the forward and the reverse pass.
#include <math.h>
#include <string.h>

#include <vector>

#include "surf_internal.cuh"

#define AS MLP_AS
#define SM_A 0
#define SM_PE (SM_A + 160 * AS)
#define SM_GPE (SM_PE + 32 * AS)
#define SM_RING (SM_GPE + 28 * AS)
#define SM_BIAS (SM_RING + MLP_NSLOT * MLP_SLOT)
#define SM_W6 (SM_BIAS + 6 * 128)
#define SM_PX (SM_W6 + 160)
#define SM_ID (SM_PX + 3 * 128)
#define SM_RED (SM_ID + 128)
#define SM_GOUT (SM_RED + 256)
#define SM_TOTAL (SM_GOUT + 3 * 128)

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

struct MlpCtx {
  float* sm;
  const DevNet* net;
  int tid, tx, ty;
  int nch;     // chunks in the stream of this launch
  int g;       // global chunk counter of this CTA
};

__device__ __forceinline__ void mlp_issue(const MlpCtx& c, int gi) {
  const int id = gi % c.nch;
  const float* src = c.net->blob + c.net->stream.off[id];
  float* dst = c.sm + SM_RING + (gi % MLP_NSLOT) * MLP_SLOT;
  const int n4 = c.net->stream.len[id] >> 2;
  for (int i = c.tid; i < n4; i += MLP_THREADS) cp_async16(dst + i * 4, src + i * 4);
  cp_async_commit();
}

// wait for chunk c.g, make it visible, refill the slot freed by chunk c.g-1 with chunk c.g+2
__device__ __forceinline__ const float* mlp_acquire(MlpCtx& c) {
  cp_async_wait<1>();
  __syncthreads();
  mlp_issue(c, c.g + 2);
  const float* w = c.sm + SM_RING + (c.g % MLP_NSLOT) * MLP_SLOT;
  c.g++;
  return w;
}

__device__ __forceinline__ int pt_of(int ty, int i) { return (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)); }

#define FMA8(ACC, AV, WV)                 \
  ACC[0] = fmaf(AV.x, WV, ACC[0]);        \
  ACC[1] = fmaf(AV.y, WV, ACC[1]);        \
  ACC[2] = fmaf(AV.z, WV, ACC[2]);        \
  ACC[3] = fmaf(AV.w, WV, ACC[3]);

// acc[j][i]: j = output slot (n = tx + 16 j), i = point slot
__device__ __forceinline__ void gemm_chunk_fwd(const float* __restrict__ Arows, const float* __restrict__ Wc, int tx,
                                               int ty, float (&acc)[8][8]) {
#pragma unroll 4
  for (int kk = 0; kk < MLP_KCH; ++kk) {
    const float4 a0 = *reinterpret_cast<const float4*>(Arows + kk * AS + ty * 4);
    const float4 a1 = *reinterpret_cast<const float4*>(Arows + kk * AS + 64 + ty * 4);
    const float4 w0 = *reinterpret_cast<const float4*>(Wc + kk * MLP_NFWD + tx * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(Wc + kk * MLP_NFWD + 64 + tx * 4);
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      FMA8((&acc[j][0]), a0, wv[j]);
      FMA8((&acc[j][4]), a1, wv[j]);
    }
  }
}

__device__ __forceinline__ void gemm_chunk_bwd(const float* __restrict__ Arows, const float* __restrict__ Wc, int tx,
                                               int ty, float (&acc)[8][8], float (&gf)[2][8]) {
#pragma unroll 4
  for (int kk = 0; kk < MLP_KCH; ++kk) {
    const float4 a0 = *reinterpret_cast<const float4*>(Arows + kk * AS + ty * 4);
    const float4 a1 = *reinterpret_cast<const float4*>(Arows + kk * AS + 64 + ty * 4);
    const float4 w0 = *reinterpret_cast<const float4*>(Wc + kk * MLP_NBWD + tx * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(Wc + kk * MLP_NBWD + 64 + tx * 4);
    const float2 w2 = *reinterpret_cast<const float2*>(Wc + kk * MLP_NBWD + 128 + tx * 2);
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      FMA8((&acc[j][0]), a0, wv[j]);
      FMA8((&acc[j][4]), a1, wv[j]);
    }
    FMA8((&gf[0][0]), a0, w2.x);
    FMA8((&gf[0][4]), a1, w2.x);
    FMA8((&gf[1][0]), a0, w2.y);
    FMA8((&gf[1][4]), a1, w2.y);
  }
}

// softplus(beta=100, threshold=20) and its derivative sigmoid(100 z)  (sdf_network.py:93)
__device__ __forceinline__ void softplus100(float z, float& h, float& dh) {
  const float t = 100.0f * z;
  if (t > 20.0f) {
    h = z;
    dh = 1.0f;
  } else {
    const float e = __expf(-fabsf(t));
    const float inv = __fdividef(1.0f, 1.0f + e);
    h = fmaxf(z, 0.f) + 0.01f * __logf(1.0f + e);
    dh = (z >= 0.f) ? inv : e * inv;
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(MLP_THREADS, 1)
k_sdf_mlp(const DevScene sc, const DevNet net, const PointSource src, float* __restrict__ sdf_out,
          float* __restrict__ grad_out, float* __restrict__ scratch_all, int negate) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  int64_t n_total = src.n;
  if (src.count) {
    const int64_t c = *src.count;
    n_total = c < n_total ? c : n_total;
  }
  const int64_t n_tiles = (n_total + MLP_TILE - 1) / MLP_TILE;
  if ((int64_t)blockIdx.x >= n_tiles) return;

  MlpCtx ctx;
  ctx.sm = sm;
  ctx.net = &net;
  ctx.tid = tid;
  ctx.tx = tx;
  ctx.ty = ty;
  ctx.nch = GRAD ? net.stream.n_chunks_all : net.stream.n_chunks_fwd;
  ctx.g = 0;
  mlp_issue(ctx, 0);
  mlp_issue(ctx, 1);

  float* A = sm + SM_A;
  float* PEb = sm + SM_PE;
  float* GPE = sm + SM_GPE;
  float* sbias = sm + SM_BIAS;
  float* sw6 = sm + SM_W6;
  float* spx = sm + SM_PX;
  int* sid = reinterpret_cast<int*>(sm + SM_ID);
  float* red = sm + SM_RED;
  float* gout = sm + SM_GOUT;
  float4* scratch = reinterpret_cast<float4*>(scratch_all) + (size_t)blockIdx.x * (6 * 16 * MLP_THREADS);

  for (int i = tid; i < 6 * 128; i += MLP_THREADS) sbias[i] = net.bias[i];
  for (int i = tid; i < 160; i += MLP_THREADS) sw6[i] = net.w6[i];
  for (int i = tid; i < 4 * AS; i += MLP_THREADS) A[156 * AS + i] = 0.f;    // K padding rows 156..159
  for (int i = tid; i < 5 * AS; i += MLP_THREADS) PEb[27 * AS + i] = 0.f;   // K padding rows 27..31
  const int pe_dim = net.pe_dim;
  const int skip = net.skip_layer;
  const int n_skip_h = MLP_HID - pe_dim;   // 101: hidden width feeding the skip layer

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    // ---- tile points -------------------------------------------------------------------------
    if (tid < MLP_TILE) {
      const int64_t i = tile * MLP_TILE + tid;
      float px = 0.f, py = 0.f, pz = 0.f;
      int64_t id = -1;
      if (i < n_total) {
        id = src.list ? (int64_t)src.list[i] : i;
        if (src.mode == 0) {
          px = src.pts[id * 3]; py = src.pts[id * 3 + 1]; pz = src.pts[id * 3 + 2];
        } else if (src.mode == 1) {
          const int64_t r = id / src.S;
          const float t = src.mid_z[id];
          px = ray_at(src.rays_o[r * 3], src.rays_d[r * 3], t);
          py = ray_at(src.rays_o[r * 3 + 1], src.rays_d[r * 3 + 1], t);
          pz = ray_at(src.rays_o[r * 3 + 2], src.rays_d[r * 3 + 2], t);
        } else {
          const int64_t yz = (int64_t)src.ny * src.nz;
          const int xi = (int)(id / yz);
          const int rem = (int)(id - (int64_t)xi * yz);
          px = src.xs[xi]; py = src.ys[rem / src.nz]; pz = src.zs[rem % src.nz];
        }
      }
      spx[tid] = px; spx[128 + tid] = py; spx[256 + tid] = pz;
      sid[tid] = (int)id;
    }
    __syncthreads();
    // ---- sparse trilinear gather -> A rows 128..155 ---------------------------------------------
    for (int it = tid; it < 4 * MLP_TILE; it += MLP_THREADS) {
      const int pt = it & (MLP_TILE - 1), l = it >> 7;
      float f[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (l < sc.n_levels) sparse_level<0>(sc, l, spx[pt], spx[128 + pt], spx[256 + pt], nullptr, f);
#pragma unroll
      for (int c = 0; c < 7; ++c) A[(128 + l * 7 + c) * AS + pt] = f[c];
    }
    // ---- positional encoding -> PEb rows 0..26 ---------------------------------------------------
    for (int it = tid; it < 3 * MLP_TILE; it += MLP_THREADS) {
      const int pt = it & (MLP_TILE - 1), d = it >> 7;
      const float x = spx[d * 128 + pt] * net.scale;
      PEb[d * AS + pt] = x;
      float fr = 1.0f;
      for (int f = 0; f < net.multires; ++f) {
        float s, c;
        sincosf(x * fr, &s, &c);
        PEb[(3 + 6 * f + d) * AS + pt] = s;
        PEb[(3 + 6 * f + 3 + d) * AS + pt] = c;
        fr *= 2.0f;
      }
      if (GRAD) {
        for (int k = d; k < 28; k += 3) GPE[k * AS + pt] = 0.f;
      }
    }

    float acc[8][8];
    // ---- forward: lin0..lin5 ----------------------------------------------------------------------
    for (int l = 0; l < 6; ++l) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
      const int nchunk = (l == 0) ? 1 : 5;
      for (int c = 0; c < nchunk; ++c) {
        const float* Wc = mlp_acquire(ctx);
        gemm_chunk_fwd((l == 0 ? PEb : A) + c * MLP_KCH * AS, Wc, tx, ty, acc);
      }
      __syncthreads();   // everybody is done reading A
      const bool to_skip = (l + 1 == skip);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = tx + 16 * j;
        const float b = sbias[l * 128 + n];
        float h[8], dh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) softplus100(acc[j][i] + b, h[i], dh[i]);
        if (to_skip && n >= n_skip_h) {   // columns 101..127 of the skip layer's input are the PE
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            h[i] = PEb[(n - n_skip_h) * AS + pt_of(ty, i)];
            dh[i] = 0.f;
          }
        }
        *reinterpret_cast<float4*>(A + n * AS + ty * 4) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(A + n * AS + 64 + ty * 4) = make_float4(h[4], h[5], h[6], h[7]);
        if (GRAD) {
          scratch[(size_t)((l * 16 + j * 2) * MLP_THREADS) + tid] = make_float4(dh[0], dh[1], dh[2], dh[3]);
          scratch[(size_t)((l * 16 + j * 2 + 1) * MLP_THREADS) + tid] = make_float4(dh[4], dh[5], dh[6], dh[7]);
        }
      }
    }
    // ---- lin6 row 0: the SDF head ------------------------------------------------------------------
    __syncthreads();
    {
      const int pt = tid & 127, half = tid >> 7;
      float s = 0.f;
#pragma unroll 8
      for (int k = half * 80; k < half * 80 + 80; ++k) s = fmaf(A[k * AS + pt], sw6[k], s);
      red[tid] = s;
    }
    __syncthreads();
    if (tid < MLP_TILE && sid[tid] >= 0) {
      const float v = (red[tid] + red[tid + 128] + net.b6) * net.inv_scale;
      sdf_out[sid[tid]] = negate ? -v : v;
    }
    if (!GRAD) continue;

    // ---- reverse pass -----------------------------------------------------------------------------
    float gf[2][8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float w = sw6[128 + tx + 16 * q] * net.inv_scale;
#pragma unroll
      for (int i = 0; i < 8; ++i) gf[q][i] = w;
    }
    // delta5 = w6[n] / scale * softplus'(z5)   (A rows 0..127 were last read by the head, synced above)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = tx + 16 * j;
      const float w = sw6[n] * net.inv_scale;
      const float4 d0 = scratch[(size_t)((5 * 16 + j * 2) * MLP_THREADS) + tid];
      const float4 d1 = scratch[(size_t)((5 * 16 + j * 2 + 1) * MLP_THREADS) + tid];
      *reinterpret_cast<float4*>(A + n * AS + ty * 4) = make_float4(w * d0.x, w * d0.y, w * d0.z, w * d0.w);
      *reinterpret_cast<float4*>(A + n * AS + 64 + ty * 4) = make_float4(w * d1.x, w * d1.y, w * d1.z, w * d1.w);
    }
    for (int l = 5; l >= 1; --l) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
      for (int c = 0; c < 4; ++c) {
        const float* Wc = mlp_acquire(ctx);
        gemm_chunk_bwd(A + c * MLP_KCH * AS, Wc, tx, ty, acc, gf);
      }
      __syncthreads();
      const bool is_skip = (l == skip);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = tx + 16 * j;
        const float4 d0 = scratch[(size_t)(((l - 1) * 16 + j * 2) * MLP_THREADS) + tid];
        const float4 d1 = scratch[(size_t)(((l - 1) * 16 + j * 2 + 1) * MLP_THREADS) + tid];
        if (is_skip && n >= n_skip_h) {
#pragma unroll
          for (int i = 0; i < 8; ++i) GPE[(n - n_skip_h) * AS + pt_of(ty, i)] = acc[j][i];
        }
        *reinterpret_cast<float4*>(A + n * AS + ty * 4) =
            make_float4(acc[j][0] * d0.x, acc[j][1] * d0.y, acc[j][2] * d0.z, acc[j][3] * d0.w);
        *reinterpret_cast<float4*>(A + n * AS + 64 + ty * 4) =
            make_float4(acc[j][4] * d1.x, acc[j][5] * d1.y, acc[j][6] * d1.z, acc[j][7] * d1.w);
      }
    }
    // ---- lin0 reverse: g_pe += delta0 . W0   (K = 128, N = 27 -> one 128 x 32 chunk) ---------------
    {
      const float* Wc = mlp_acquire(ctx);
      const int pt = tid & 127, half = tid >> 7;
      float a16[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) a16[q] = 0.f;
#pragma unroll 2
      for (int k = 0; k < MLP_HID; ++k) {
        const float a = A[k * AS + pt];
        const float4* w4 = reinterpret_cast<const float4*>(Wc + k * 32 + half * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = w4[q];
          a16[q * 4 + 0] = fmaf(a, w.x, a16[q * 4 + 0]);
          a16[q * 4 + 1] = fmaf(a, w.y, a16[q * 4 + 1]);
          a16[q * 4 + 2] = fmaf(a, w.z, a16[q * 4 + 2]);
          a16[q * 4 + 3] = fmaf(a, w.w, a16[q * 4 + 3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int n = half * 16 + q;
        if (n < pe_dim) GPE[n * AS + pt] += a16[q];
      }
    }
    __syncthreads();
    // feature gradients (registers) -> A rows 128..155
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n = 128 + tx + 16 * q;
      if (n < 156) {
        *reinterpret_cast<float4*>(A + n * AS + ty * 4) = make_float4(gf[q][0], gf[q][1], gf[q][2], gf[q][3]);
        *reinterpret_cast<float4*>(A + n * AS + 64 + ty * 4) = make_float4(gf[q][4], gf[q][5], gf[q][6], gf[q][7]);
      }
    }
    __syncthreads();
    // d PE / d x
    for (int it = tid; it < 3 * MLP_TILE; it += MLP_THREADS) {
      const int pt = it & (MLP_TILE - 1), d = it >> 7;
      float gx = GPE[d * AS + pt];
      float fr = 1.0f;
      for (int f = 0; f < net.multires; ++f) {
        const float s = PEb[(3 + 6 * f + d) * AS + pt], c = PEb[(3 + 6 * f + 3 + d) * AS + pt];
        gx += fr * (GPE[(3 + 6 * f + d) * AS + pt] * c - GPE[(3 + 6 * f + 3 + d) * AS + pt] * s);
        fr *= 2.0f;
      }
      gout[d * 128 + pt] = gx * net.scale;
    }
    // d feats / d x, one (point, level) per thread; partials into A rows 0..11
    for (int it = tid; it < 4 * MLP_TILE; it += MLP_THREADS) {
      const int pt = it & (MLP_TILE - 1), l = it >> 7;
      float o3[3] = {0.f, 0.f, 0.f};
      if (l < sc.n_levels) {
        float g7[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) g7[c] = A[(128 + l * 7 + c) * AS + pt];
        sparse_level<1>(sc, l, spx[pt], spx[128 + pt], spx[256 + pt], g7, o3);
      }
      A[(l * 3 + 0) * AS + pt] = o3[0];
      A[(l * 3 + 1) * AS + pt] = o3[1];
      A[(l * 3 + 2) * AS + pt] = o3[2];
    }
    __syncthreads();
    if (tid < MLP_TILE && sid[tid] >= 0) {
      const int64_t id = sid[tid];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float g = gout[d * 128 + tid];
#pragma unroll
        for (int l = 0; l < 4; ++l) g += A[(l * 3 + d) * AS + tid];
        grad_out[id * 3 + d] = g;
      }
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Full (n, d_out) output of SDFNetworkSparse.forward (sdf_network.py:95-121): [sdf / scale, lin6 rows 1..].
// The 128 extra outputs are dead on the render path (implicit_surface.py:95-97), so the hot kernels above only
// evaluate row 0; this plain fp32 kernel exists for API completeness (`sdf_network(x, volumes, indexes)`).
// 16 points per 128-thread block, thread = one output neuron, activations in shared memory, weights k-major in
// global memory (coalesced, L1/L2 resident: 480 KB).
// ---------------------------------------------------------------------------------------------
#define FH_PTS 16
#define FH_NS 160          // row stride of the k-major weight matrices
__global__ void __launch_bounds__(128)
k_sdf_full(const DevScene sc, const DevNet net, const float* __restrict__ wfull, const int* __restrict__ woff,
           const float* __restrict__ pts, int64_t n, int d_out, float* __restrict__ out) {
  __shared__ float s_in[FH_PTS][160];
  __shared__ float s_out[FH_PTS][160];
  __shared__ float s_pe[FH_PTS][28];
  __shared__ float s_ft[FH_PTS][28];
  const int tid = threadIdx.x;
  for (int64_t base = (int64_t)blockIdx.x * FH_PTS; base < n; base += (int64_t)gridDim.x * FH_PTS) {
    __syncthreads();
    // features (thread = (point, level)) and positional encoding (thread = point)
    if (tid < FH_PTS * 4) {
      const int p = tid >> 2, lv = tid & 3;
      const int64_t i = base + p;
      float f7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (i < n && lv < sc.n_levels) sparse_level<0>(sc, lv, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], nullptr, f7);
#pragma unroll
      for (int c = 0; c < 7; ++c) s_ft[p][lv * 7 + c] = f7[c];
    } else if (tid < FH_PTS * 5) {
      const int p = tid - FH_PTS * 4;
      const int64_t i = base + p;
      float x[3] = {0.f, 0.f, 0.f};
      if (i < n) { x[0] = pts[i * 3] * net.scale; x[1] = pts[i * 3 + 1] * net.scale; x[2] = pts[i * 3 + 2] * net.scale; }
      for (int d = 0; d < 3; ++d) s_pe[p][d] = x[d];
      float fr = 1.0f;
      for (int f = 0; f < net.multires; ++f) {
        for (int d = 0; d < 3; ++d) {
          s_pe[p][3 + 6 * f + d] = sinf(x[d] * fr);
          s_pe[p][3 + 6 * f + 3 + d] = cosf(x[d] * fr);
        }
        fr *= 2.0f;
      }
    }
    __syncthreads();
    for (int i = tid; i < FH_PTS * net.pe_dim; i += 128) s_in[i / net.pe_dim][i % net.pe_dim] = s_pe[i / net.pe_dim][i % net.pe_dim];
    __syncthreads();
    int in_dim = net.pe_dim;
    for (int l = 0; l < 7; ++l) {
      const int od = (l == 6) ? d_out : net.out_dim[l];
      const float* Wl = wfull + woff[l];
      for (int nb = 0; nb < od; nb += 128) {
        const int o = nb + tid;
        float acc[FH_PTS];
#pragma unroll
        for (int p = 0; p < FH_PTS; ++p) acc[p] = 0.f;
        if (o < od) {
          for (int k = 0; k < in_dim; ++k) {
            const float w = Wl[(size_t)k * FH_NS + o];
#pragma unroll
            for (int p = 0; p < FH_PTS; ++p) acc[p] = fmaf(s_in[p][k], w, acc[p]);
          }
          const float b = Wl[(size_t)in_dim * FH_NS + o];        // bias row
#pragma unroll
          for (int p = 0; p < FH_PTS; ++p) {
            float v = acc[p] + b;
            if (l < 6) {
              float h, dh;
              softplus100(v, h, dh);
              s_out[p][o] = h;
            } else {
              const int64_t i = base + p;
              if (i < n) out[i * d_out + o] = (o == 0) ? v * net.inv_scale : v;
            }
          }
        }
      }
      if (l == 6) break;
      __syncthreads();
      // next input: [h (out_dim) | PE if the next layer is the skip layer | feats]; the 1/sqrt(2) of the skip concat
      // is folded into the weights
      int nd = od;
      for (int i = tid; i < FH_PTS * od; i += 128) s_in[i / od][i % od] = s_out[i / od][i % od];
      if (l + 1 == net.skip_layer) {
        for (int i = tid; i < FH_PTS * net.pe_dim; i += 128) s_in[i / net.pe_dim][od + i % net.pe_dim] = s_pe[i / net.pe_dim][i % net.pe_dim];
        nd += net.pe_dim;
      }
      for (int i = tid; i < FH_PTS * 28; i += 128) s_in[i / 28][nd + i % 28] = s_ft[i / 28][i % 28];
      in_dim = nd + 28;
      __syncthreads();
    }
  }
}

// grid pre-pass for the opt-in sparsified extraction (Q16): mask -> list, fill elsewhere
__global__ void k_grid_sparsify(const DevScene sc, const float* __restrict__ xs, const float* __restrict__ ys,
                                const float* __restrict__ zs, int nx, int ny, int nz, float fill,
                                float* __restrict__ u, int32_t* __restrict__ list, int32_t* __restrict__ counter) {
  const int64_t n = (int64_t)nx * ny * nz;
  const int lane = threadIdx.x & 31;
  for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n;
       base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = base + lane;
    bool valid = false;
    if (id < n) {
      const int64_t yz = (int64_t)ny * nz;
      const int xi = (int)(id / yz);
      const int rem = (int)(id - (int64_t)xi * yz);
      valid = scene_point_mask(sc, xs[xi], ys[rem / nz], zs[rem % nz]);
      if (!valid) u[id] = fill;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (bal) {
      int start = 0;
      if (lane == 0) start = atomicAdd(counter, __popc(bal));
      start = __shfl_sync(0xffffffffu, start, 0);
      if (valid) list[start + __popc(bal & ((1u << lane) - 1))] = (int32_t)id;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host: weight folding + stream construction
// ---------------------------------------------------------------------------------------------
static inline int perm_pos_128(int n) {   // column n -> position inside a permuted 128-wide row
  const int tx = n & 15, jj = n >> 4;     // n = tx + 16 * jj, jj = 4 h + q
  return 64 * (jj >> 2) + tx * 4 + (jj & 3);
}

int surf_build_smooth_weights(const std::vector<std::vector<float>>& W, const surf_net_inputs* in, surf_net* net,
                              cudaStream_t st, int (*dev_alloc)(surf_net*, void**, size_t));
int surf_build_tc_weights(const std::vector<std::vector<float>>& W, const surf_net_inputs* in, surf_net* net,
                          cudaStream_t st, int (*dev_alloc)(surf_net*, void**, size_t));
int surf_build_smooth_tc_weights(const std::vector<std::vector<float>>& W, const surf_net_inputs* in, surf_net* net,
                                 cudaStream_t st, int (*dev_alloc)(surf_net*, void**, size_t));

int surf_build_sdf_weights(const surf_net_inputs* in, surf_net* net, cudaStream_t st,
                           int (*dev_alloc)(surf_net*, void**, size_t)) {
  SURF_CHECK_ARG(in->n_lin == SURF_SDF_LAYERS, "n_lin must be 7");
  SURF_CHECK_ARG(in->feat_channels == 28, "feat_channels must be 28 (4 levels x 7)");
  const int pe_dim = in->multires > 0 ? 3 + 6 * in->multires : 3;
  SURF_CHECK_ARG(pe_dim <= 27, "multires must be <= 4");
  SURF_CHECK_ARG(in->in_dim[0] == pe_dim, "lin0 input dim must equal the PE dim");
  const int skip = in->skip_layer;
  for (int l = 1; l < 7; ++l) SURF_CHECK_ARG(in->in_dim[l] == 156, "hidden input dim must be 156");
  for (int l = 0; l < 6; ++l) {
    const int want = (l + 1 == skip) ? 128 - pe_dim : 128;
    SURF_CHECK_ARG(in->out_dim[l] == want, "hidden output dims must be 128 (128 - PE before the skip layer)");
  }
  SURF_CHECK_ARG(in->out_dim[6] >= 1, "lin6 out dim");
  // fold weight norm: W = g * v / ||v||_row  (nn.utils.weight_norm dim=0; sdf_network.py:88-89)
  std::vector<std::vector<float>> W(7);
  for (int l = 0; l < 7; ++l) {
    const int O = in->out_dim[l], I = in->in_dim[l];
    SURF_CHECK_ARG(in->h_weight_v[l] && in->h_bias[l], "null weight pointer");
    W[l].resize((size_t)O * I);
    for (int o = 0; o < O; ++o) {
      const float* v = in->h_weight_v[l] + (size_t)o * I;
      float sc = 1.0f;
      if (in->h_weight_g[l]) {
        double ss = 0.0;
        for (int i = 0; i < I; ++i) ss += (double)v[i] * v[i];
        sc = in->h_weight_g[l][o] / (float)sqrt(ss);
      }
      for (int i = 0; i < I; ++i) W[l][(size_t)o * I + i] = v[i] * sc;
    }
  }
  if (skip >= 1 && skip <= 5) {   // fold the 1/sqrt(2) of the skip concat into the first 128 input columns
    const float r = (float)(1.0 / sqrt(2.0));
    const int O = in->out_dim[skip], I = in->in_dim[skip];
    for (int o = 0; o < O; ++o)
      for (int i = 0; i < 128; ++i) W[skip][(size_t)o * I + i] *= r;
  }
  MlpStream& s = net->dev.stream;
  memset(&s, 0, sizeof(s));
  std::vector<float> blob;
  int nc = 0;
  auto add_chunk = [&](size_t len) {
    s.off[nc] = (int)blob.size();
    s.len[nc] = (int)len;
    blob.resize(blob.size() + len, 0.f);
    return blob.size() - len;
  };
  // forward chunks: rows k (input), 128 permuted output columns
  for (int l = 0; l < 6; ++l) {
    const int O = in->out_dim[l], I = in->in_dim[l];
    const int nchunk = (l == 0) ? 1 : 5;
    for (int c = 0; c < nchunk; ++c) {
      const size_t base = add_chunk((size_t)MLP_KCH * MLP_NFWD);
      nc++;
      for (int kk = 0; kk < MLP_KCH; ++kk) {
        const int k = c * MLP_KCH + kk;
        if (k >= I) continue;
        for (int n = 0; n < O && n < 128; ++n) blob[base + (size_t)kk * MLP_NFWD + perm_pos_128(n)] = W[l][(size_t)n * I + k];
      }
    }
  }
  s.n_chunks_fwd = nc;
  // reverse chunks of lin5..lin1: rows k (output index), 160 permuted input columns
  for (int l = 5; l >= 1; --l) {
    const int O = in->out_dim[l], I = in->in_dim[l];
    for (int c = 0; c < 4; ++c) {
      const size_t base = add_chunk((size_t)MLP_KCH * MLP_NBWD);
      nc++;
      for (int kk = 0; kk < MLP_KCH; ++kk) {
        const int k = c * MLP_KCH + kk;
        if (k >= O) continue;
        for (int n = 0; n < I; ++n) {
          const int pos = n < 128 ? perm_pos_128(n) : 128 + ((n - 128) & 15) * 2 + ((n - 128) >> 4);
          blob[base + (size_t)kk * MLP_NBWD + pos] = W[l][(size_t)k * I + n];
        }
      }
    }
  }
  {  // reverse of lin0: 128 x 32, natural order
    const int O = in->out_dim[0], I = in->in_dim[0];
    const size_t base = add_chunk((size_t)128 * 32);
    nc++;
    for (int k = 0; k < O && k < 128; ++k)
      for (int n = 0; n < I; ++n) blob[base + (size_t)k * 32 + n] = W[0][(size_t)k * I + n];
  }
  s.n_chunks_all = nc;
  if (nc > MLP_MAXCHUNK) {
    surf_set_error("weight stream too long");
    return -1;
  }
  std::vector<float> bias(6 * 128, 0.f), w6(160, 0.f);
  for (int l = 0; l < 6; ++l)
    for (int n = 0; n < in->out_dim[l]; ++n) bias[l * 128 + n] = in->h_bias[l][n];
  for (int k = 0; k < in->in_dim[6]; ++k) w6[k] = W[6][k];
  void* p = nullptr;
  int rc = dev_alloc(net, &p, blob.size() * sizeof(float));
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  net->dev.blob = (const float*)p;
  rc = dev_alloc(net, &p, bias.size() * sizeof(float));
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  net->dev.bias = (const float*)p;
  rc = dev_alloc(net, &p, w6.size() * sizeof(float));
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, w6.data(), w6.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  net->dev.w6 = (const float*)p;
  {   // k-major fp32 matrices (+ bias row) of all 7 layers for the full-head kernel k_sdf_full
    std::vector<float> wf;
    std::vector<int> off(8, 0);
    for (int l = 0; l < 7; ++l) {
      const int O = in->out_dim[l], I = in->in_dim[l];
      if (O > FH_NS) {
        surf_set_error("lin%d: out dim %d > %d unsupported", l, O, FH_NS);
        return -1;
      }
      off[l] = (int)wf.size();
      wf.resize(wf.size() + (size_t)(I + 1) * FH_NS, 0.f);
      for (int k = 0; k < I; ++k)
        for (int o = 0; o < O; ++o) wf[off[l] + (size_t)k * FH_NS + o] = W[l][(size_t)o * I + k];
      for (int o = 0; o < O; ++o) wf[off[l] + (size_t)I * FH_NS + o] = in->h_bias[l][o];
    }
    rc = dev_alloc(net, &p, wf.size() * sizeof(float));
    if (rc) return rc;
    SURF_CUDA(cudaMemcpyAsync(p, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    net->w_full = (const float*)p;
    rc = dev_alloc(net, &p, off.size() * sizeof(int));
    if (rc) return rc;
    SURF_CUDA(cudaMemcpyAsync(p, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    net->w_full_off = (const int*)p;
    SURF_CUDA(cudaStreamSynchronize(st));
  }
  rc = surf_build_smooth_weights(W, in, net, st, dev_alloc);
  if (rc) return rc;
  SURF_CUDA(cudaStreamSynchronize(st));   // host vectors go out of scope
  net->tc_ok = (in->multires == 4 && skip == 3) ? 1 : 0;
  if (net->tc_ok) {
    rc = surf_build_tc_weights(W, in, net, st, dev_alloc);
    if (rc) return rc;
    rc = surf_build_smooth_tc_weights(W, in, net, st, dev_alloc);
    if (rc) return rc;
  }
  net->dev.b6 = in->h_bias[6][0];
  net->dev.scale = in->scale;
  net->dev.inv_scale = 1.0f / in->scale;
  net->dev.multires = in->multires;
  net->dev.pe_dim = pe_dim;
  net->dev.skip_layer = skip;
  net->dev.n_feat = in->feat_channels;
  for (int l = 0; l < 7; ++l) net->dev.out_dim[l] = in->out_dim[l];
  return 0;
}

int launch_sdf_mlp(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_sdf, float* d_grad,
                   bool negate, int mode, cudaStream_t st) {
  if (src.n <= 0) return 0;
  SURF_CHECK_ARG(mode == SURF_MLP_FFMA || mode == SURF_MLP_TC || mode == SURF_MLP_TC_FAST, "mlp_mode must be SURF_MLP_FFMA, SURF_MLP_TC or SURF_MLP_TC_FAST");
  if (mode != SURF_MLP_FFMA && n->tc_ok) return launch_sdf_tc2(s, n, src, d_sdf, d_grad, negate, mode == SURF_MLP_TC_FAST, st);
  const size_t smem = SM_TOTAL * sizeof(float);
  int rc = surf_ensure_dyn_smem((const void*)k_sdf_mlp<true>, (int)smem);
  if (rc) return rc;
  rc = surf_ensure_dyn_smem((const void*)k_sdf_mlp<false>, (int)smem);
  if (rc) return rc;
  int64_t tiles = (src.n + MLP_TILE - 1) / MLP_TILE;
  int grid = (int)(tiles < n->n_sm ? tiles : n->n_sm);
  surf_time_begin(d_grad ? 0 : 1, st);
  if (d_grad) {
    k_sdf_mlp<true><<<grid, MLP_THREADS, smem, st>>>(s->dev, n->dev, src, d_sdf, d_grad, n->scratch, negate ? 1 : 0);
  } else {
    k_sdf_mlp<false><<<grid, MLP_THREADS, smem, st>>>(s->dev, n->dev, src, d_sdf, nullptr, n->scratch, negate ? 1 : 0);
  }
  surf_time_end(d_grad ? 0 : 1, st);
  SURF_LAUNCH_CHECK();
  return 0;
}

extern "C" int surf_sdf_points(const surf_scene* s, const surf_net* n, const float* d_pts, int64_t n_pts,
                               float* d_sdf, float* d_grad, int32_t mlp_mode, void* stream) {
  if (n_pts <= 0) return 0;
  SURF_CHECK_ARG(s && n && d_pts && d_sdf, "null pointer");
  SURF_CHECK_ARG(n_pts < 0x7fffffffll, "too many points for one call");
  PointSource src;
  memset(&src, 0, sizeof(src));
  src.mode = 0;
  src.pts = d_pts;
  src.n = n_pts;
  return launch_sdf_mlp(s, n, src, d_sdf, d_grad, false, mlp_mode, (cudaStream_t)stream);
}

extern "C" int surf_sdf_full(const surf_scene* s, const surf_net* n, const float* d_pts, int64_t n_pts, float* d_out,
                             int32_t d_out_dim, void* stream) {
  if (n_pts <= 0) return 0;
  SURF_CHECK_ARG(s && n && d_pts && d_out, "null pointer");
  SURF_CHECK_ARG(d_out_dim == n->dev.out_dim[6], "d_out_dim must equal the out dim of lin6");
  const int64_t blocks = (n_pts + FH_PTS - 1) / FH_PTS;
  const int64_t cap = (int64_t)surf_num_sms() * 16;
  k_sdf_full<<<(int)(blocks < cap ? blocks : cap), 128, 0, (cudaStream_t)stream>>>(s->dev, n->dev, n->w_full, n->w_full_off,
                                                                                  d_pts, n_pts, d_out_dim, d_out);
  SURF_LAUNCH_CHECK();
  return 0;
}

extern "C" int surf_sdf_grid(const surf_scene* s, const surf_net* n, const float* d_xs, int32_t nx, const float* d_ys,
                             int32_t ny, const float* d_zs, int32_t nz, float* d_u, int32_t sparsify, float fill,
                             int32_t mlp_mode, void* stream) {
  SURF_CHECK_ARG(s && n && d_xs && d_ys && d_zs && d_u, "null pointer");
  SURF_CHECK_ARG(nx > 0 && ny > 0 && nz > 0, "empty grid");
  const int64_t total = (int64_t)nx * ny * nz;
  SURF_CHECK_ARG(total < 0x7fffffffll, "grid too large for one call (shard it)");
  cudaStream_t st = (cudaStream_t)stream;
  PointSource src;
  memset(&src, 0, sizeof(src));
  src.mode = 2;
  src.xs = d_xs; src.ys = d_ys; src.zs = d_zs;
  src.nx = nx; src.ny = ny; src.nz = nz;
  src.n = total;
  int32_t* list = nullptr;
  if (sparsify) {
    SURF_CUDA(cudaMallocAsync((void**)&list, (size_t)(total + 1) * sizeof(int32_t), st));
    int32_t* counter = list + total;
    SURF_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
    int64_t g = (total + 255) / 256;
    const int64_t cap = (int64_t)surf_num_sms() * 8;
    k_grid_sparsify<<<(int)(g < cap ? g : cap), 256, 0, st>>>(s->dev, d_xs, d_ys, d_zs, nx, ny, nz, fill, d_u, list,
                                                               counter);
    SURF_LAUNCH_CHECK();
    src.list = list;
    src.count = counter;
  }
  int rc = launch_sdf_mlp(s, n, src, d_u, nullptr, true, mlp_mode, st);
  if (list) cudaFreeAsync(list, st);
  return rc;
}
