// K2b: projection of sample points into the source views + bilinear feature / RGB gather
//      (lookup_feature + compute_angle, projector.py:485-556, quirk Q14)
// K3b: IBRNet-style colour blending MLP (BlendingNetwork.forward, blending_network.py:69-117)
//
// K2b is an HBM/L2-bound gather: per (point, view) 4 taps x 32 B (RGB + level-0 features fused in one
// texel) + 3 levels x 4 taps x 16 B, all 128-bit loads from NHWC maps.  K3b (fp32 FFMA edition) runs
// 128-row (point, view) tiles through register-tiled shared-memory GEMMs, weights resident in smem.
#include <math.h>
#include <string.h>

#include <vector>

#include "lookup_common.cuh"
#include "surf_internal.cuh"

// one thread per (point, source view)
__global__ void __launch_bounds__(256, 4)
k_lookup_feature(const DevScene sc, const PointSource src, float* __restrict__ feat_out, float* __restrict__ rd_out,
                 uint8_t* __restrict__ mask_out, int packed19) {
  int64_t n_total = src.n;
  if (src.count) {
    const int64_t c = *src.count;
    n_total = c < n_total ? c : n_total;
  }
  const int V = sc.V;
  const int64_t work = n_total * V;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < work; it += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = it / V;
    const int v = (int)(it - i * V);
    float px, py, pz;
    if (src.mode == 0) {
      px = src.pts[i * 3]; py = src.pts[i * 3 + 1]; pz = src.pts[i * 3 + 2];
    } else {
      const int64_t id = src.list ? (int64_t)src.list[i] : i;
      const int64_t r = id / src.S;
      const float t = src.mid_z[id];
      px = ray_at(src.rays_o[r * 3], src.rays_d[r * 3], t);
      py = ray_at(src.rays_o[r * 3 + 1], src.rays_d[r * 3 + 1], t);
      pz = ray_at(src.rays_o[r * 3 + 2], src.rays_d[r * 3 + 2], t);
    }
    float rec[FEAT_REC];
    float4 rd;
    lookup_row(sc, px, py, pz, v, rec, rd);
    const bool valid = rec[19] > 0.f;
    if (packed19) {
      float* o = feat_out + it * 19;
#pragma unroll
      for (int c = 0; c < 19; ++c) o[c] = rec[c];
      if (mask_out) mask_out[it] = valid ? 1 : 0;
    } else {
      float4* o = reinterpret_cast<float4*>(feat_out + it * FEAT_REC);
#pragma unroll
      for (int c = 0; c < 5; ++c) o[c] = make_float4(rec[4 * c], rec[4 * c + 1], rec[4 * c + 2], rec[4 * c + 3]);
    }
    reinterpret_cast<float4*>(rd_out)[it] = rd;
  }
}

// ---------------------------------------------------------------------------------------------
// blending MLP: tiles of 128 (point, view) rows, activations k-major in shared memory, every Linear
// a small register-tiled GEMM (thread = 4 rows x N/8 outputs), weights resident in shared memory.
// ---------------------------------------------------------------------------------------------
#define BL_THREADS 256
#define BL_ROWS 128
#define BAS 132   // row stride of the k-major activation buffers

enum { L_DIR0, L_DIR1, L_BASE0, L_BASE1, L_VIS0, L_VIS1, L_V20, L_V21, L_RGB0, L_RGB1, L_RGB2, L_COUNT };
//                         K   NJ (N = 8 NJ)
static const int h_blend_K[L_COUNT] = {4, 16, 57, 64, 32, 32, 32, 32, 37, 16, 8};
static const int h_blend_NJ[L_COUNT] = {2, 3, 8, 4, 4, 5, 4, 1, 2, 1, 1};
static const int h_blend_out[L_COUNT] = {16, 19, 64, 32, 32, 33, 32, 1, 16, 8, 1};

struct BlendOffsets { int w[L_COUNT]; int b[L_COUNT]; int total; };
static BlendOffsets blend_offsets() {
  BlendOffsets o;
  int off = 0;
  for (int l = 0; l < L_COUNT; ++l) {
    o.w[l] = off;
    off += h_blend_K[l] * 8 * h_blend_NJ[l];
    o.b[l] = off;
    off += 8 * h_blend_NJ[l];
    off = (off + 3) & ~3;
  }
  o.total = off;
  return o;
}
__constant__ BlendOffsets c_blend_off;

// smem layout (floats)
#define BS_RD 0
#define BS_F (BS_RD + 4 * BAS)
#define BS_XA (BS_F + 20 * BAS)
#define BS_XB (BS_XA + 64 * BAS)
#define BS_WV (BS_XB + 64 * BAS)
#define BS_M (BS_WV + BL_ROWS)
#define BS_VIS (BS_M + BL_ROWS)
#define BS_LOGIT (BS_VIS + BL_ROWS)
#define BS_W (BS_LOGIT + BL_ROWS)

__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : (__expf(x) - 1.0f); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// acc[j][i] = sum_k X[k][ty*4+i] * W'[k][tx*NJ + j]   (output n = tx + 8 j)
template <int NJ>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ X, const float* __restrict__ Wp, int K, int tx,
                                          int ty, float (&acc)[NJ][4]) {
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(X + k * BAS + ty * 4);
    const float* w = Wp + k * (8 * NJ) + tx * NJ;
    float wv[NJ];
    if (NJ % 4 == 0) {
#pragma unroll
      for (int q = 0; q < NJ / 4; ++q) {
        const float4 t = reinterpret_cast<const float4*>(w)[q];
        wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NJ; ++j) wv[j] = w[j];
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      acc[j][0] = fmaf(a.x, wv[j], acc[j][0]);
      acc[j][1] = fmaf(a.y, wv[j], acc[j][1]);
      acc[j][2] = fmaf(a.z, wv[j], acc[j][2]);
      acc[j][3] = fmaf(a.w, wv[j], acc[j][3]);
    }
  }
}

__device__ __forceinline__ void store4(float* X, int n, int ty, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(X + n * BAS + ty * 4) = make_float4(a, b, c, d);
}

__global__ void __launch_bounds__(BL_THREADS, 1)
k_blend(const float* __restrict__ blob, float s_abs, const float* __restrict__ feat, const float* __restrict__ rdiff,
        const uint8_t* __restrict__ mask, int V, int packed19, const int32_t* __restrict__ list,
        const int32_t* __restrict__ count, int64_t n, float* __restrict__ rgb_out, uint8_t* __restrict__ views_out) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  float* RD = sm + BS_RD;
  float* F = sm + BS_F;
  float* XA = sm + BS_XA;
  float* XB = sm + BS_XB;
  float* s_wv = sm + BS_WV;
  float* s_m = sm + BS_M;
  float* s_vis = sm + BS_VIS;
  float* s_logit = sm + BS_LOGIT;
  float* W = sm + BS_W;
  for (int i = tid; i < c_blend_off.total; i += BL_THREADS) W[i] = blob[i];
  int64_t n_total = n;
  if (count) {
    const int64_t c = *count;
    n_total = c < n_total ? c : n_total;
  }
  const int ppt = BL_ROWS / V;          // points per tile
  const int rows = ppt * V;
  const int64_t n_tiles = (n_total + ppt - 1) / ppt;
#define WL(l) (W + c_blend_off.w[l])
#define BL(l) (W + c_blend_off.b[l])
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    // ---- load records: row r <-> (point tile*ppt + r / V, view r % V) ----
    if (tid < BL_ROWS) {
      const int r = tid;
      const int64_t rec_i = tile * rows + r;
      const bool ok = (r < rows) && (rec_i < n_total * V);
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
      for (int c = 0; c < FEAT_REC; ++c) F[c * BAS + r] = rec[c];
      RD[0 * BAS + r] = rd.x; RD[1 * BAS + r] = rd.y; RD[2 * BAS + r] = rd.z; RD[3 * BAS + r] = rd.w;
      s_m[r] = rec[19];
    }
    __syncthreads();
    // ---- anti-alias pooling weights (blending_network.py:76-80) ----
    // weight = (e_v - min_v e_v) * mask / (sum + 1e-8) with e_v = exp(|s| (dot_v - 1)): a difference of
    // nearly equal exponentials.  The reference's result is quantised to fp32 ulps of e_v, so e_v is
    // formed here as the correctly rounded fp32 exponential (fp64 exp, rounded once) of the identically
    // rounded fp32 argument: that keeps this path within the reference's own 1-ulp envelope.
    if (tid < BL_ROWS) {
      const float arg = __fmul_rn(s_abs, __fsub_rn(RD[3 * BAS + tid], 1.0f));
      s_logit[tid] = (float)exp((double)arg);     // s_logit doubles as scratch for e_v until the last layer
    }
    __syncthreads();
    if (tid < ppt) {
      const int r0 = tid * V;
      float emin = INFINITY;
      for (int v = 0; v < V; ++v) emin = fminf(emin, s_logit[r0 + v]);
      float wsum = 0.f;
      for (int v = 0; v < V; ++v) wsum = __fadd_rn(wsum, __fmul_rn(__fsub_rn(s_logit[r0 + v], emin), s_m[r0 + v]));
      const float den = __fadd_rn(wsum, 1e-8f);
      for (int v = 0; v < V; ++v)
        s_wv[r0 + v] = __fdiv_rn(__fmul_rn(__fsub_rn(s_logit[r0 + v], emin), s_m[r0 + v]), den);
    }
    // ---- ray_dir_fc: 4 -> 16 -> 19, ELU; x = feat + dir (kept in XB rows 38..56) ----
    {
      float acc[2][4];
      tile_gemm<2>(RD, WL(L_DIR0), 4, tx, ty, acc);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int nn = tx + 8 * j;
        const float b = BL(L_DIR0)[nn];
        store4(XA, nn, ty, elu(acc[j][0] + b), elu(acc[j][1] + b), elu(acc[j][2] + b), elu(acc[j][3] + b));
      }
    }
    __syncthreads();
    {
      float acc[3][4];
      tile_gemm<3>(XA, WL(L_DIR1), 16, tx, ty, acc);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int nn = tx + 8 * j;
        if (nn < 19) {
          const float b = BL(L_DIR1)[nn];
          const float4 f = *reinterpret_cast<const float4*>(F + nn * BAS + ty * 4);
          store4(XB, 38 + nn, ty, f.x + elu(acc[j][0] + b), f.y + elu(acc[j][1] + b), f.z + elu(acc[j][2] + b),
                 f.w + elu(acc[j][3] + b));
        }
      }
    }
    __syncthreads();
    // ---- weighted mean / variance over views (fused_mean_variance, :15-19) -> XB rows 0..37 ----
    for (int it = tid; it < 19 * ppt; it += BL_THREADS) {
      const int c = it / ppt, pp = it - c * ppt;
      const int r0 = pp * V;
      const float* x = XB + (38 + c) * BAS + r0;
      float mean = 0.f;
      for (int v = 0; v < V; ++v) mean = fmaf(x[v], s_wv[r0 + v], mean);
      float var = 0.f;
      for (int v = 0; v < V; ++v) {
        const float d = x[v] - mean;
        var = fmaf(s_wv[r0 + v] * d, d, var);
      }
      for (int v = 0; v < V; ++v) {
        XB[c * BAS + r0 + v] = mean;
        XB[(19 + c) * BAS + r0 + v] = var;
      }
    }
    __syncthreads();
    // ---- base_fc: 57 -> 64 -> 32 ----
    {
      float acc[8][4];
      tile_gemm<8>(XB, WL(L_BASE0), 57, tx, ty, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int nn = tx + 8 * j;
        const float b = BL(L_BASE0)[nn];
        store4(XA, nn, ty, elu(acc[j][0] + b), elu(acc[j][1] + b), elu(acc[j][2] + b), elu(acc[j][3] + b));
      }
    }
    __syncthreads();
    {
      float acc[4][4];
      tile_gemm<4>(XA, WL(L_BASE1), 64, tx, ty, acc);
      __syncthreads();   // XB (base_fc input) no longer read by anyone
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = tx + 8 * j;
        const float b = BL(L_BASE1)[nn];
        store4(XB, nn, ty, elu(acc[j][0] + b), elu(acc[j][1] + b), elu(acc[j][2] + b), elu(acc[j][3] + b));
      }
    }
    __syncthreads();
    // ---- vis_fc(x * weight): the per-row scalar commutes with the Linear ----
    {
      float acc[4][4];
      tile_gemm<4>(XB, WL(L_VIS0), 32, tx, ty, acc);
      const float4 wv = *reinterpret_cast<const float4*>(s_wv + ty * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = tx + 8 * j;
        const float b = BL(L_VIS0)[nn];
        store4(XA, nn, ty, elu(acc[j][0] * wv.x + b), elu(acc[j][1] * wv.y + b), elu(acc[j][2] * wv.z + b),
               elu(acc[j][3] * wv.w + b));
      }
    }
    __syncthreads();
    {
      float acc[5][4];
      tile_gemm<5>(XA, WL(L_VIS1), 32, tx, ty, acc);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int nn = tx + 8 * j;
        const float b = BL(L_VIS1)[nn];
        if (nn < 32) {          // x = x + x_res
          float4 x = *reinterpret_cast<const float4*>(XB + nn * BAS + ty * 4);
          x.x += elu(acc[j][0] + b); x.y += elu(acc[j][1] + b); x.z += elu(acc[j][2] + b); x.w += elu(acc[j][3] + b);
          *reinterpret_cast<float4*>(XB + nn * BAS + ty * 4) = x;
        } else if (nn == 32) {  // vis = sigmoid(vis) * mask
#pragma unroll
          for (int i = 0; i < 4; ++i) s_vis[ty * 4 + i] = sigmoidf_(elu(acc[j][i] + b)) * s_m[ty * 4 + i];
        }
      }
    }
    __syncthreads();
    // ---- vis_fc2(x * vis) * mask ----
    {
      float acc[4][4];
      tile_gemm<4>(XB, WL(L_V20), 32, tx, ty, acc);
      const float4 vs = *reinterpret_cast<const float4*>(s_vis + ty * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = tx + 8 * j;
        const float b = BL(L_V20)[nn];
        store4(XA, nn, ty, elu(acc[j][0] * vs.x + b), elu(acc[j][1] * vs.y + b), elu(acc[j][2] * vs.z + b),
               elu(acc[j][3] * vs.w + b));
      }
    }
    __syncthreads();
    {
      float acc[1][4];
      tile_gemm<1>(XA, WL(L_V21), 32, tx, ty, acc);
      if (tx == 0) {
        const float b = BL(L_V21)[0];
        store4(XB, 32, ty, sigmoidf_(acc[0][0] + b) * s_m[ty * 4], sigmoidf_(acc[0][1] + b) * s_m[ty * 4 + 1],
               sigmoidf_(acc[0][2] + b) * s_m[ty * 4 + 2], sigmoidf_(acc[0][3] + b) * s_m[ty * 4 + 3]);
      } else if (tx <= 4) {   // rows 33..36 <- ray_diff
        *reinterpret_cast<float4*>(XB + (32 + tx) * BAS + ty * 4) =
            *reinterpret_cast<const float4*>(RD + (tx - 1) * BAS + ty * 4);
      }
    }
    __syncthreads();
    // ---- rgb_fc: 37 -> 16 -> 8 -> 1 ----
    {
      float acc[2][4];
      tile_gemm<2>(XB, WL(L_RGB0), 37, tx, ty, acc);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int nn = tx + 8 * j;
        const float b = BL(L_RGB0)[nn];
        store4(XA, nn, ty, elu(acc[j][0] + b), elu(acc[j][1] + b), elu(acc[j][2] + b), elu(acc[j][3] + b));
      }
    }
    __syncthreads();
    {
      float acc[1][4];
      tile_gemm<1>(XA, WL(L_RGB1), 16, tx, ty, acc);
      const float b = BL(L_RGB1)[tx];
      store4(XB, tx, ty, elu(acc[0][0] + b), elu(acc[0][1] + b), elu(acc[0][2] + b), elu(acc[0][3] + b));
    }
    __syncthreads();
    {
      float acc[1][4];
      tile_gemm<1>(XB, WL(L_RGB2), 8, tx, ty, acc);
      if (tx == 0) {
        const float b = BL(L_RGB2)[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) s_logit[ty * 4 + i] = acc[0][i] + b;
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
          cr = fmaf(e, F[0 * BAS + r0 + v], cr);
          cg = fmaf(e, F[1 * BAS + r0 + v], cg);
          cb = fmaf(e, F[2 * BAS + r0 + v], cb);
        }
        const int64_t id = list ? (int64_t)list[i] : i;
        const float inv = 1.0f / ssum;
        rgb_out[id * 3] = cr * inv;
        rgb_out[id * 3 + 1] = cg * inv;
        rgb_out[id * 3 + 2] = cb * inv;
        if (views_out) views_out[id] = (uint8_t)vbits;
      }
    }
  }
#undef WL
#undef BL
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
// W (out,in) row-major -> W'[k][tx*NJ + j] with n = tx + 8 j ; bias natural order
static void put_layer(std::vector<float>& blob, const BlendOffsets& o, int l, const float* Wm, const float* b, int in_dim) {
  const int K = h_blend_K[l], NJ = h_blend_NJ[l], out = h_blend_out[l];
  for (int k = 0; k < K && k < in_dim; ++k)
    for (int n = 0; n < out; ++n) {
      const int tx = n & 7, j = n >> 3;
      blob[o.w[l] + k * 8 * NJ + tx * NJ + j] = Wm[(size_t)n * in_dim + k];
    }
  for (int n = 0; n < out; ++n) blob[o.b[l] + n] = b[n];
}

int surf_build_blend_weights(const surf_net_inputs* in, surf_net* net, cudaStream_t st,
                             int (*dev_alloc)(surf_net*, void**, size_t)) {
  SURF_CHECK_ARG(in->d_feature == 16, "d_feature must be 16");
  for (int i = 0; i < 11; ++i) SURF_CHECK_ARG(in->h_blend_w[i] && in->h_blend_b[i], "null blend weight");
  const BlendOffsets o = blend_offsets();
  std::vector<float> blob(o.total, 0.f);
  for (int l = 0; l < L_COUNT; ++l) put_layer(blob, o, l, in->h_blend_w[l], in->h_blend_b[l], h_blend_K[l]);
  void* p = nullptr;
  int rc = dev_alloc(net, &p, blob.size() * sizeof(float));
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  SURF_CUDA(cudaMemcpyToSymbolAsync(c_blend_off, &o, sizeof(o), 0, cudaMemcpyHostToDevice, st));
  SURF_CUDA(cudaStreamSynchronize(st));
  net->dev.blend = (const float*)p;
  net->dev.blend_s = fabsf(in->blend_s);
  return 0;
}

static int cap_blocks(int64_t n, int per_block, int per_sm) {
  int64_t g = (n + per_block - 1) / per_block;
  const int64_t cap = (int64_t)surf_num_sms() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int launch_lookup_feature(const surf_scene* s, const PointSource& src, float* d_feat, float* d_raydiff,
                          uint8_t* d_mask, bool packed19, cudaStream_t st, bool small_blocks) {
  SURF_CHECK_ARG(s->dev.img0, "scene has no images / feature maps");
  if (src.n <= 0 || s->dev.V <= 0) return 0;
  surf_time_begin(2, st);
  if (small_blocks)     // 128 threads x 64 registers: fits in the registers the persistent SDF kernel leaves free
    k_lookup_feature<<<cap_blocks(src.n * s->dev.V, 128, 16), 128, 0, st>>>(s->dev, src, d_feat, d_raydiff, d_mask,
                                                                            packed19 ? 1 : 0);
  else
    k_lookup_feature<<<cap_blocks(src.n * s->dev.V, 256, 8), 256, 0, st>>>(s->dev, src, d_feat, d_raydiff, d_mask,
                                                                           packed19 ? 1 : 0);
  surf_time_end(2, st);
  SURF_LAUNCH_CHECK();
  return 0;
}

int launch_blend(const surf_scene*, const surf_net* n, const float* d_feat, const float* d_raydiff,
                 const uint8_t* d_mask, int V, bool packed19, const int32_t* list, const int32_t* count, int64_t n_pts,
                 float* d_rgb, uint8_t* d_views, int mode, cudaStream_t st) {
  if (n_pts <= 0) return 0;
  SURF_CHECK_ARG(V >= 1 && V <= SURF_MAX_VIEWS, "n_src_views");
  SURF_CHECK_ARG(mode == SURF_MLP_FFMA || mode == SURF_MLP_TC || mode == SURF_MLP_TC_FAST, "mlp_mode must be SURF_MLP_FFMA, SURF_MLP_TC or SURF_MLP_TC_FAST");
  if (mode != SURF_MLP_FFMA)
    return launch_blend_tc(n, d_feat, d_raydiff, d_mask, V, packed19, list, count, n_pts, d_rgb, d_views,
                           mode == SURF_MLP_TC_FAST, st);
  const size_t smem = (size_t)(BS_W + blend_offsets().total) * sizeof(float);
  int rc = surf_ensure_dyn_smem((const void*)k_blend, (int)smem);
  if (rc) return rc;
  const int ppt = BL_ROWS / V;
  int64_t tiles = (n_pts + ppt - 1) / ppt;
  const int grid = (int)(tiles < n->n_sm ? tiles : n->n_sm);
  surf_time_begin(3, st);
  k_blend<<<grid, BL_THREADS, smem, st>>>(n->dev.blend, n->dev.blend_s, d_feat, d_raydiff, d_mask, V,
                                          packed19 ? 1 : 0, list, count, n_pts, d_rgb, d_views);
  surf_time_end(3, st);
  SURF_LAUNCH_CHECK();
  return 0;
}

extern "C" int surf_lookup_feature(const surf_scene* s, const float* d_pts, int64_t n_pts, float* d_feat_views,
                                   float* d_ray_diff, uint8_t* d_mask, void* stream) {
  SURF_CHECK_ARG(s && d_pts && d_feat_views && d_ray_diff && d_mask, "null pointer");
  PointSource src;
  memset(&src, 0, sizeof(src));
  src.mode = 0;
  src.pts = d_pts;
  src.n = n_pts;
  return launch_lookup_feature(s, src, d_feat_views, d_ray_diff, d_mask, true, (cudaStream_t)stream);
}

extern "C" int surf_blend(const surf_net* n, const float* d_feat_views, const float* d_ray_diff, const uint8_t* d_mask,
                          int64_t n_pts, int32_t n_src_views, float* d_rgb, int32_t mlp_mode, void* stream) {
  SURF_CHECK_ARG(n && d_feat_views && d_ray_diff && d_mask && d_rgb, "null pointer");
  SURF_CHECK_ARG(n_src_views >= 1 && n_src_views <= SURF_MAX_VIEWS, "n_src_views");
  return launch_blend(nullptr, n, d_feat_views, d_ray_diff, d_mask, n_src_views, true, nullptr, nullptr, n_pts, d_rgb,
                      nullptr, mlp_mode, (cudaStream_t)stream);
}
