// Producers of the scene tensors (models/modules/volume.py:54-168).
//   k_back_proj        : Volume.back_proj_multiscale — one thread per voxel: project into every view, sum the bilinear
//                        samples (grid_sample, align_corners=True, zeros) of the feature scales stage.., score each view
//                        with the 4 -> 8 -> 1 ELU MLP, masked softmax over the views -> (mean | "variance") (n, 2c),
//                        and the >= 2-view frustum mask.
//   k_depth_filter     : Volume.depth_filtering — a voxel survives when its camera depth agrees with the rendered depth
//                        map (bilinear, align_corners=True) within depth_range in at least two views.
//   k_upsample3d_2x    : F.interpolate(scale_factor=2, mode="trilinear") of the previous matching volume (sparse2dense).
//   k_scatter_*        : sparse2dense / get_index emitting the COMPACT layout of the render path directly: int32 index
//                        table, 1-bit masks, fp32 matching volume — the reference's fp32 mask volumes and int64 tables
//                        (6 GB at 704^3, 45 GB at 1408^3) never exist (surf_scene_create_sparse).
// HBM/L2-bound gathers and scatters, once per scene and stage.
#include "surf_internal.cuh"

struct BackProjViews {
  int nv;
  float w2c[SURF_MAXV + 1][12];     // rows 0..2 of inverse(c2w)
  float K[SURF_MAXV + 1][12];       // rows 0..2 of the 4x4 intrinsics
};
struct BackProjFeats {
  int n_scales, c;
  const float* f[4];                // (nv, c, h, w) NCHW, the scales that are summed
  int h[4], w[4];
  int hn, wn;                       // size of the FINEST map: the normalisation of the projection (volume.py:74-75)
};
struct AggMlp {
  float w0[8][4], b0[8], w1[8], b1;
};

__device__ __forceinline__ bool vol_project(const BackProjViews& V, int v, float wx, float wy, float wz, int hn, int wn,
                                            float* nx, float* ny, float* depth) {
  const float* M = V.w2c[v];
  const float cx = M[0] * wx + M[1] * wy + M[2] * wz + M[3];
  const float cy = M[4] * wx + M[5] * wy + M[6] * wz + M[7];
  const float cz = M[8] * wx + M[9] * wy + M[10] * wz + M[11];
  const float* K = V.K[v];
  const float ix = K[0] * cx + K[1] * cy + K[2] * cz + K[3];
  const float iy = K[4] * cx + K[5] * cy + K[6] * cz + K[7];
  const float iz = K[8] * cx + K[9] * cy + K[10] * cz + K[11];
  const float x = __fdiv_rn(ix, iz), y = __fdiv_rn(iy, iz);
  *nx = __fsub_rn(__fdiv_rn(x, (float)(wn - 1) * 0.5f), 1.0f);
  *ny = __fsub_rn(__fdiv_rn(y, (float)(hn - 1) * 0.5f), 1.0f);
  *depth = iz;
  return fabsf(*nx) <= 1.0f && fabsf(*ny) <= 1.0f && iz > 0.f;
}

// grid_sample(bilinear, zeros, align_corners=True) of one NCHW channel plane set at normalised (nx, ny)
__device__ __forceinline__ void vol_sample(const float* __restrict__ f, int c, int h, int w, float nx, float ny, float* acc) {
  const float ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.0f), 0.5f), (float)(w - 1));
  const float iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.0f), 0.5f), (float)(h - 1));
  if (!(ix > -1.0f && ix < (float)w && iy > -1.0f && iy < (float)h)) return;      // (also NaN)
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix, wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int xi = x0 + (t & 1), yi = y0 + (t >> 1);
    if (xi < 0 || xi >= w || yi < 0 || yi >= h) continue;
    const float wt = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0);
    for (int ch = 0; ch < c; ++ch) acc[ch] += __ldg(f + ((size_t)ch * h + yi) * w + xi) * wt;
  }
}

__global__ void __launch_bounds__(256)
k_back_proj(const BackProjViews V, const BackProjFeats F, const AggMlp A, const float* __restrict__ coords, int64_t n,
            float3 vs, float3 org, float* __restrict__ feat_vol, uint8_t* __restrict__ mask_vol) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float wx = __fadd_rn(__fmul_rn(coords[i * 3], vs.x), org.x);
    const float wy = __fadd_rn(__fmul_rn(coords[i * 3 + 1], vs.y), org.y);
    const float wz = __fadd_rn(__fmul_rn(coords[i * 3 + 2], vs.z), org.z);
    float warp[SURF_MAXV + 1][4], logit[SURF_MAXV + 1];
    int seen = 0;
    float mx = -INFINITY;
    for (int v = 0; v < V.nv; ++v) {
      float nx, ny, dz;
      const bool ok = vol_project(V, v, wx, wy, wz, F.hn, F.wn, &nx, &ny, &dz);
      seen += ok ? 1 : 0;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int s = 0; s < F.n_scales; ++s)
        vol_sample(F.f[s] + (size_t)v * F.c * F.h[s] * F.w[s], F.c, F.h[s], F.w[s], nx, ny, acc);
      float lg = A.b1;
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        float a = A.b0[o];
#pragma unroll
        for (int k = 0; k < 4; ++k) a = fmaf(A.w0[o][k], acc[k], a);
        a = a > 0.f ? a : expm1f(a);
        lg = fmaf(A.w1[o], a, lg);
      }
      if (!ok) lg = -1e9f;
#pragma unroll
      for (int k = 0; k < 4; ++k) warp[v][k] = acc[k];
      logit[v] = lg;
      mx = fmaxf(mx, lg);
    }
    float ssum = 0.f;
    for (int v = 0; v < V.nv; ++v) { logit[v] = expf(logit[v] - mx); ssum += logit[v]; }
    float mean[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int v = 0; v < V.nv; ++v) {
      const float wv = logit[v] / ssum;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float t = warp[v][k] * wv;
        mean[k] += t;
        sq[k] += t * t;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      feat_vol[i * 8 + k] = mean[k];
      feat_vol[i * 8 + 4 + k] = sq[k] - mean[k] * mean[k];          // the reference's "variance" (volume.py:91)
    }
    mask_vol[i] = seen > 1 ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256)
k_depth_filter(const BackProjViews V, const float* __restrict__ depths, int h, int w, const float* __restrict__ coords,
               int64_t n, float3 vs, float3 org, float depth_range, uint8_t* __restrict__ valid) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float wx = __fadd_rn(__fmul_rn(coords[i * 3], vs.x), org.x);
    const float wy = __fadd_rn(__fmul_rn(coords[i * 3 + 1], vs.y), org.y);
    const float wz = __fadd_rn(__fmul_rn(coords[i * 3 + 2], vs.z), org.z);
    int cnt = 0;
    for (int v = 0; v < V.nv; ++v) {
      float nx, ny, dz;
      const bool ok = vol_project(V, v, wx, wy, wz, h, w, &nx, &ny, &dz);
      float d = 0.f;
      vol_sample(depths + (size_t)v * h * w, 1, h, w, nx, ny, &d);
      cnt += (ok && fabsf(__fsub_rn(d, dz)) < depth_range) ? 1 : 0;
    }
    valid[i] = cnt > 1 ? 1 : 0;
  }
}

// F.interpolate(scale_factor=2, mode="trilinear", align_corners=False): (D,H,W) -> (2D,2H,2W)
__global__ void __launch_bounds__(256)
k_upsample3d_2x(const float* __restrict__ in, int D, int H, int W, float* __restrict__ out) {
  const int oD = 2 * D, oH = 2 * H, oW = 2 * W;
  const int64_t n = (int64_t)oD * oH * oW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % oW), y = (int)((i / oW) % oH), z = (int)(i / ((int64_t)oW * oH));
    int z0, z1, y0, y1, x0, x1;
    float lz, ly, lx;
    up_src(z, 0.5f, D, &z0, &z1, &lz);
    up_src(y, 0.5f, H, &y0, &y1, &ly);
    up_src(x, 0.5f, W, &x0, &x1, &lx);
    const float hz = __fsub_rn(1.0f, lz), hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
    auto at = [&](int zz, int yy, int xx) { return __ldg(in + ((size_t)zz * H + yy) * W + xx); };
    auto row = [&](int zz, int yy) { return __fadd_rn(__fmul_rn(hx, at(zz, yy, x0)), __fmul_rn(lx, at(zz, yy, x1))); };
    auto plane = [&](int zz) { return __fadd_rn(__fmul_rn(hy, row(zz, y0)), __fmul_rn(ly, row(zz, y1))); };
    out[i] = __fadd_rn(__fmul_rn(hz, plane(z0)), __fmul_rn(lz, plane(z1)));
  }
}

__global__ void k_fill_i32(int32_t* __restrict__ p, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
// sparse2dense + get_index in the compact layout: row id into the int32 table, mask bit, matching logit
__global__ void k_scatter_level(const float* __restrict__ coords, int64_t n, int N, int32_t* __restrict__ index,
                                uint32_t* __restrict__ bits, const float* __restrict__ logits, float* __restrict__ matching) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)coords[i * 3], y = (int)coords[i * 3 + 1], z = (int)coords[i * 3 + 2];
    if (x < 0 || x >= N || y < 0 || y >= N || z < 0 || z >= N) continue;
    const size_t lin = ((size_t)x * N + y) * N + z;
    index[lin] = (int32_t)i;
    atomicOr(bits + (lin >> 5), 1u << (lin & 31));
    if (matching && logits) matching[lin] = logits[i];
  }
}

static void fill_views(BackProjViews* V, int nv, const float* h_w2cs, const float* h_intrs) {
  V->nv = nv;
  for (int v = 0; v < nv; ++v)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        V->w2c[v][r * 4 + c] = h_w2cs[(size_t)v * 16 + r * 4 + c];
        V->K[v][r * 4 + c] = h_intrs[(size_t)v * 16 + r * 4 + c];
      }
}
static inline int vol_grid(int64_t n, int threads) {
  const int64_t b = (n + threads - 1) / threads, cap = (int64_t)surf_num_sms() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

extern "C" int surf_volume_back_proj(const surf_volume_views* vw, const float* const* d_feats, const int32_t* feat_h,
                                     const int32_t* feat_w, int32_t n_scales, int32_t n_channels, const float* h_agg_mlp,
                                     const float* d_coords, int64_t n_vox, const float* voxel_size, const float* origin,
                                     float* d_feat_vol, uint8_t* d_mask_vol, void* stream) {
  if (n_vox <= 0) return 0;
  SURF_CHECK_ARG(vw && d_feats && feat_h && feat_w && h_agg_mlp && d_coords && voxel_size && origin && d_feat_vol && d_mask_vol, "null pointer");
  SURF_CHECK_ARG(vw->n_views >= 1 && vw->n_views <= SURF_MAXV + 1, "n_views out of range");
  SURF_CHECK_ARG(n_scales >= 1 && n_scales <= 4 && n_channels == 4, "1..4 feature scales of 4 channels (agg_mlp is 4 -> 8 -> 1)");
  BackProjViews V;
  fill_views(&V, vw->n_views, vw->h_w2cs, vw->h_intrs);
  BackProjFeats F;
  F.n_scales = n_scales; F.c = n_channels;
  for (int s = 0; s < n_scales; ++s) { F.f[s] = d_feats[s]; F.h[s] = feat_h[s]; F.w[s] = feat_w[s]; }
  F.hn = vw->norm_h; F.wn = vw->norm_w;
  AggMlp A;
  const float* p = h_agg_mlp;                 // [0.weight (8,4) | 0.bias (8) | 2.weight (1,8) | 2.bias (1)]
  for (int o = 0; o < 8; ++o) for (int k = 0; k < 4; ++k) A.w0[o][k] = p[o * 4 + k];
  for (int o = 0; o < 8; ++o) A.b0[o] = p[32 + o];
  for (int o = 0; o < 8; ++o) A.w1[o] = p[40 + o];
  A.b1 = p[48];
  k_back_proj<<<vol_grid(n_vox, 256), 256, 0, (cudaStream_t)stream>>>(V, F, A, d_coords, n_vox,
      make_float3(voxel_size[0], voxel_size[1], voxel_size[2]), make_float3(origin[0], origin[1], origin[2]), d_feat_vol, d_mask_vol);
  SURF_LAUNCH_CHECK();
  return 0;
}

extern "C" int surf_volume_depth_filter(const surf_volume_views* vw, const float* d_depths, const float* d_coords,
                                        int64_t n_vox, const float* voxel_size, const float* origin, float depth_range,
                                        uint8_t* d_valid, void* stream) {
  if (n_vox <= 0) return 0;
  SURF_CHECK_ARG(vw && d_depths && d_coords && voxel_size && origin && d_valid, "null pointer");
  SURF_CHECK_ARG(vw->n_views >= 1 && vw->n_views <= SURF_MAXV + 1, "n_views out of range");
  BackProjViews V;
  fill_views(&V, vw->n_views, vw->h_w2cs, vw->h_intrs);
  k_depth_filter<<<vol_grid(n_vox, 256), 256, 0, (cudaStream_t)stream>>>(V, d_depths, vw->norm_h, vw->norm_w, d_coords, n_vox,
      make_float3(voxel_size[0], voxel_size[1], voxel_size[2]), make_float3(origin[0], origin[1], origin[2]), depth_range, d_valid);
  SURF_LAUNCH_CHECK();
  return 0;
}

extern "C" int surf_volume_upsample2x(const float* d_in, int32_t D, int32_t H, int32_t W, float* d_out, void* stream) {
  SURF_CHECK_ARG(d_in && d_out && D >= 1 && H >= 1 && W >= 1, "bad arguments");
  k_upsample3d_2x<<<vol_grid((int64_t)8 * D * H * W, 256), 256, 0, (cudaStream_t)stream>>>(d_in, D, H, W, d_out);
  SURF_LAUNCH_CHECK();
  return 0;
}

int scene_alloc_pub(surf_scene* s, void** p, size_t bytes);
int scene_pad_volume_pub(surf_scene* s, int level, const float* d_volume, int64_t n_vox, cudaStream_t st);

// build_volumes' outputs (surf.py:80-131) straight into the prepared layout: per level the voxel coordinates (rows of
// volumes[l] in the same order), the 7-channel features and the matching logits; levels COARSE -> FINE as they are built.
extern "C" int surf_scene_create_sparse(const surf_scene_sparse_inputs* in, void* stream, surf_scene** out) {
  SURF_CHECK_ARG(in && out, "inputs/out null");
  SURF_CHECK_ARG(in->n_levels >= 1 && in->n_levels <= SURF_MAX_LEVELS, "n_levels must be 1..4");
  SURF_CHECK_ARG(in->feat_ch >= 1 && in->feat_ch <= 7, "feat_ch must be 1..7");
  cudaStream_t st = (cudaStream_t)stream;
  surf_scene* s = new surf_scene();
  memset(s, 0, sizeof(*s));
  DevScene& d = s->dev;
  d.n_levels = in->n_levels;
  d.feat_ch = in->feat_ch;
  int rc = 0;
  float* prev_match = nullptr;
  int prev_dim = 0;
  const int L = in->n_levels;
  for (int b = 0; b < L; ++b) {               // b: build order (coarse -> fine); renderer level l = L - 1 - b
    const int l = L - 1 - b;
    const int N = in->dim[b];
    const int64_t nvx = in->n_vox[b];
    if (N < 2 || !in->d_coords[b] || (!in->d_volumes[b] && nvx > 0) || (b > 0 && N != 2 * prev_dim)) {
      surf_set_error("level %d: bad dim / null tensor (dims must double from level to level)", b);
      surf_scene_destroy(s);
      return -1;
    }
    d.dim[l] = N;
    d.voxel[l] = 2.0f / (float)(N - 1);
    s->nvox[l] = nvx;
    const size_t n3 = (size_t)N * N * N;
    void* p = nullptr;
    if ((rc = scene_alloc_pub(s, &p, n3 * sizeof(int32_t)))) { surf_scene_destroy(s); return rc; }
    int32_t* index = (int32_t*)p;
    d.index[l] = index;
    const size_t n_words = (n3 + 31) / 32;
    if ((rc = scene_alloc_pub(s, &p, n_words * sizeof(uint32_t)))) { surf_scene_destroy(s); return rc; }
    uint32_t* bits = (uint32_t*)p;
    d.mask[l] = bits;
    float* match = nullptr;
    if (in->d_logits[b]) {
      if ((rc = scene_alloc_pub(s, &p, n3 * sizeof(float)))) { surf_scene_destroy(s); return rc; }
      match = (float*)p;
      if (prev_match) {
        k_upsample3d_2x<<<vol_grid((int64_t)n3, 256), 256, 0, st>>>(prev_match, prev_dim, prev_dim, prev_dim, match);
        surf_count_launch();
      } else {
        cudaMemsetAsync(match, 0, n3 * sizeof(float), st);
      }
    }
    k_fill_i32<<<vol_grid((int64_t)n3, 256), 256, 0, st>>>(index, (int64_t)n3, -1);
    cudaMemsetAsync(bits, 0, n_words * sizeof(uint32_t), st);
    if (nvx > 0) {
      k_scatter_level<<<vol_grid(nvx, 256), 256, 0, st>>>(in->d_coords[b], nvx, N, index, bits, in->d_logits[b], match);
      surf_count_launch(2);
      if ((rc = scene_pad_volume_pub(s, l, in->d_volumes[b], nvx, st))) { surf_scene_destroy(s); return rc; }
    }
    s->stats.bytes_index += n3 * sizeof(int32_t);
    s->stats.bytes_masks += n_words * sizeof(uint32_t);
    s->stats.n_vox[l] = nvx;
    prev_match = match;
    prev_dim = N;
  }
  if (prev_match) {           // the finest level's volume is THE matching volume (volume.py:105-110, surf.py:115)
    d.matching = prev_match;
    d.mdim = prev_dim;
    s->stats.bytes_matching = (size_t)prev_dim * prev_dim * prev_dim * sizeof(float);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    surf_set_error("scene_create_sparse: %s", cudaGetErrorString(e));
    surf_scene_destroy(s);
    return (int)e;
  }
  *out = s;
  return 0;
}
