// Internal declarations shared by the surf_b200 translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "surf_b200.h"

#define SURF_MAXV SURF_MAX_VIEWS

// ---------------------------------------------------------------------------------------------
// device-side scene view (POD, passed to kernels by value)
// ---------------------------------------------------------------------------------------------
struct DevScene {
  int n_levels, feat_ch;
  int dim[SURF_MAX_LEVELS];
  float voxel[SURF_MAX_LEVELS];          // fp32 2/(N-1), computed like torch (projector.py:231)
  const int32_t* index[SURF_MAX_LEVELS];  // N^3 int32, -1 = empty
  const float4* vol8[SURF_MAX_LEVELS];    // nvox rows of 8 floats (7 used) = 2 float4 = one 32 B sector
  const uint32_t* mask[SURF_MAX_LEVELS];  // N^3 bits
  const float* matching;                  // M^3 fp32
  int mdim;
  int nv, V, H, W;                        // V = nv-1 source views
  int fh[4], fw[4];                       // feature pyramid sizes
  const float4* img0;                     // (nv,H,W) texels of 2 float4: [r,g,b,f0a | f0b,f0c,f0d,0]
  const float4* feat[4];                  // levels 1..3: (nv,h,w) float4 ; [0] unused
  float w2c[SURF_MAXV][12];               // source view v (= view v+1): rows 0..2 of inverse(c2w)
  float K[SURF_MAXV][9];                  // intrinsics[:3,:3] of source view v (level 0)
  float cen[SURF_MAXV][3];                // source camera centres c2w[:3,3]
  float refcen[3];                        // reference camera centre
  float rot0inv[9];                       // inverse(c2w0)[:3,:3]
};

struct surf_scene {
  DevScene dev;
  void* owned[64];
  int n_owned;
  void* view_buf[4];                      // NHWC image / feature maps of the current views (surf_scene_set_views)
  size_t view_bytes[4];
  int views_version;                      // bumped by surf_scene_set_views
  void* warp12;                           // extras.cu: 12-channel warp maps of the current views, NHWC
  size_t warp_bytes;
  int warp_version;                       // views_version the maps were built from
  surf_scene_stats stats;
  int64_t nvox[SURF_MAX_LEVELS];
};

// ---------------------------------------------------------------------------------------------
// SDF MLP weight stream (see sdf_mlp.cu)
// ---------------------------------------------------------------------------------------------
#define MLP_TILE 128        // points per tile
#define MLP_THREADS 256
#define MLP_HID 128
#define MLP_KCH 32          // k rows per weight chunk
#define MLP_AS 132          // activation row stride (floats): 132 % 32 == 4 -> conflict-free column writes
#define MLP_NFWD 128        // forward N tile
#define MLP_NBWD 160        // backward N tile (128 hidden + 28 feats + 4 pad)
#define MLP_SLOT (MLP_KCH * MLP_NBWD)   // ring slot in floats (20 KB)
#define MLP_NSLOT 3
#define MLP_MAXCHUNK 64

struct MlpStream {
  int n_chunks_fwd;                 // chunks of the forward-only stream
  int n_chunks_all;                 // forward + backward
  int off[MLP_MAXCHUNK];            // float offset into blob
  int len[MLP_MAXCHUNK];            // floats (multiple of 4)
};

struct DevNet {
  const float* blob;                // weight stream
  const float* bias;                // [6][128] biases of lin0..lin5 (padded)
  const float* w6;                  // [160] row 0 of lin6 (skip-scaled not needed), padded with 0
  float b6;                         // bias[0] of lin6
  float inv_scale, scale;
  int multires, pe_dim, skip_layer, n_feat;   // 4, 27, 3, 28
  int out_dim[SURF_SDF_LAYERS];
  const float* blend;               // blend weight blob (see blend.cu for the layout)
  float blend_s;
  float inv_s;                      // clip(exp(10*variance), 1e-6, 1e6)
  MlpStream stream;
};

// weight stream of the tensor-core SDF kernel (sdf_tc2.cu): byte offsets / sizes of the fp16 hi|lo chunks
#define T1_MAXCHUNK 64
struct T1Stream {
  int n_fwd, n_all;
  uint32_t off[T1_MAXCHUNK];      // byte offset in the blob
  uint32_t bytes[T1_MAXCHUNK];
};

struct surf_net {
  DevNet dev;
  void* owned[24];
  int n_owned;
  const uint8_t* tc_blob;           // tensor-core weight stream (fp16 hi/lo chunks, forward then reverse)
  T1Stream tc_stream;               // chunk table of tc_blob in the order sdf_tc2.cu consumes it
  void* tc_scratch;                 // softplus' code scratch of sdf_tc2.cu (per-CTA private, L2-resident)
  const uint8_t* blend_tc_w;        // blend_tc.cu: fp16 hi/lo tensor-core operands
  const float* blend_tc_f;          // blend_tc.cu: fp32 small-layer weights and biases
  const float* w_full;              // k-major fp32 matrices (+ bias row) of lin0..lin6 for k_sdf_full (opt-in (n,129) head)
  const int* w_full_off;            // float offset of each layer in w_full
  const float* w_rows;              // row-major fp32 matrices (stride 160) for the reverse passes of k_sdf_smooth
  const int* w_rows_off;
  const uint8_t* smooth_tc_w;       // sdf_smooth_tc.cu: fp16 hi/lo operands of all 12 steps, in stream order
  float2* smooth_tc_scratch;        // ... and its per-CTA softplus-derivative scratch
  int tc_ok;                        // network shape supported by the tensor-core kernels
  float* scratch;                   // sigma' scratch of the FFMA backward pass (per-CTA private)
  size_t scratch_bytes;
  int n_sm;
};

// ---------------------------------------------------------------------------------------------
// error handling / launch accounting
// ---------------------------------------------------------------------------------------------
void surf_set_error(const char* fmt, ...);
// bench-only kernel timing (scene.cu): RAII-less begin/end pair around a launch
void surf_time_begin(int kind, cudaStream_t st);
void surf_time_end(int kind, cudaStream_t st);
void surf_count_launch(int n = 1);
int surf_num_sms();
int surf_ensure_dyn_smem(const void* func, int bytes);   // per (device, kernel), thread-safe

#define SURF_CHECK_ARG(cond, msg)            \
  do {                                       \
    if (!(cond)) {                           \
      surf_set_error("invalid argument: %s", msg); \
      return -1;                             \
    }                                        \
  } while (0)

#define SURF_CUDA(expr)                                                      \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      surf_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                        \
    }                                                                        \
  } while (0)

#define SURF_LAUNCH_CHECK()                                                  \
  do {                                                                       \
    surf_count_launch();                                                     \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) {                                                 \
      surf_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                        \
    }                                                                        \
  } while (0)

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// ATen grid_sampler_unnormalize, align_corners=False: ((c + 1) * size - 1) / 2 with separate
// roundings (the CPU reference is compiled without FMA contraction); bit-exact restatement.
__device__ __forceinline__ float gs_unnorm(float c, int size) {
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(c, 1.0f), (float)size), 1.0f), 0.5f);
}

// o + d * t with separate roundings (implicit_surface.py:80,286: mul then add as two ATen ops)
__device__ __forceinline__ float ray_at(float o, float d, float t) { return __fadd_rn(o, __fmul_rn(d, t)); }

// lookup_volume(..., 'nearest') on one bit-packed level (projector.py:415 -> grid_sampler_3d nearest,
// zeros padding).  World (x,y,z) <-> volume (D,H,W).
__device__ __forceinline__ bool mask_nearest(const uint32_t* __restrict__ bits, int N, float px, float py, float pz) {
  float fx = rintf(gs_unnorm(px, N));   // D index
  float fy = rintf(gs_unnorm(py, N));   // H index
  float fz = rintf(gs_unnorm(pz, N));   // W index
  if (!(fx >= 0.f && fx < (float)N && fy >= 0.f && fy < (float)N && fz >= 0.f && fz < (float)N)) return false;
  size_t lin = ((size_t)(int)fx * N + (int)fy) * N + (int)fz;
  return (bits[lin >> 5] >> (lin & 31)) & 1u;
}

__device__ __forceinline__ bool scene_point_mask(const DevScene& sc, float px, float py, float pz) {
  bool m = false;
#pragma unroll
  for (int l = 0; l < SURF_MAX_LEVELS; ++l)
    if (l < sc.n_levels) m = m || mask_nearest(sc.mask[l], sc.dim[l], px, py, pz);
  return m;
}

// Sparse trilinear fetch of one level (projector.py:217-374, Q13).
//  MODE 0: out[c] = sum_corner val[c] * w                       (c < 7)
//  MODE 1: out[0..2] = d/d(px,py,pz) of  sum_c g[c] * feat[c]   (gradient w.r.t. the world point)
template <int MODE>
__device__ __forceinline__ void sparse_level(const DevScene& sc, int l, float px, float py, float pz,
                                             const float* g, float* out) {
  const int N = sc.dim[l];
  const float vs = sc.voxel[l];
  // flipped point: grid x <- world z, grid y <- world y, grid z <- world x (projector.py:379)
  const float cx = __fdiv_rn(__fadd_rn(pz, 1.0f), vs);
  const float cy = __fdiv_rn(__fadd_rn(py, 1.0f), vs);
  const float cz = __fdiv_rn(__fadd_rn(px, 1.0f), vs);
  const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
  // weights from the UNCLAMPED corners
  const float wx1 = __fsub_rn(cx, fx0), wx0 = __fsub_rn(fx0 + 1.0f, cx);
  const float wy1 = __fsub_rn(cy, fy0), wy0 = __fsub_rn(fy0 + 1.0f, cy);
  const float wz1 = __fsub_rn(cz, fz0), wz0 = __fsub_rn(fz0 + 1.0f, cz);
  // clamp in float first so far-away points cannot overflow the int conversion
  const float hi = (float)(N - 1);
  const int x0 = (int)fminf(fmaxf(fx0, 0.f), hi), x1 = (int)fminf(fmaxf(fx0 + 1.0f, 0.f), hi);
  const int y0 = (int)fminf(fmaxf(fy0, 0.f), hi), y1 = (int)fminf(fmaxf(fy0 + 1.0f, 0.f), hi);
  const int z0 = (int)fminf(fmaxf(fz0, 0.f), hi), z1 = (int)fminf(fmaxf(fz0 + 1.0f, 0.f), hi);
  const int32_t* __restrict__ idx = sc.index[l];
  const float4* __restrict__ vol = sc.vol8[l];
  int32_t rows[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int xi = (c & 1) ? x1 : x0, yi = (c & 2) ? y1 : y0, zi = (c & 4) ? z1 : z0;
    rows[c] = __ldg(idx + ((size_t)zi * N + yi) * N + xi);
  }
  if (MODE == 0) {
#pragma unroll
    for (int c = 0; c < 7; ++c) out[c] = 0.f;
  } else {
    out[0] = out[1] = out[2] = 0.f;
  }
  float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {   // order bnw,bne,bsw,bse,fnw,fne,fsw,fse (projector.py:360-371)
    if (rows[c] < 0) continue;
    const float4 a = __ldg(vol + (size_t)rows[c] * 2);
    const float4 b = __ldg(vol + (size_t)rows[c] * 2 + 1);
    const float wx = (c & 1) ? wx1 : wx0, wy = (c & 2) ? wy1 : wy0, wz = (c & 4) ? wz1 : wz0;
    if (MODE == 0) {
      const float w = __fmul_rn(__fmul_rn(wx, wy), wz);
      out[0] += a.x * w; out[1] += a.y * w; out[2] += a.z * w; out[3] += a.w * w;
      out[4] += b.x * w; out[5] += b.y * w; out[6] += b.z * w;
    } else {
      const float s = a.x * g[0] + a.y * g[1] + a.z * g[2] + a.w * g[3] + b.x * g[4] + b.y * g[5] + b.z * g[6];
      gx += s * (((c & 1) ? 1.f : -1.f) * wy * wz);
      gy += s * (((c & 2) ? 1.f : -1.f) * wx * wz);
      gz += s * (((c & 4) ? 1.f : -1.f) * wx * wy);
    }
  }
  if (MODE == 1) {
    const float inv = 1.0f / vs;
    out[0] = gz * inv;   // d/d world x  (grid z)
    out[1] = gy * inv;
    out[2] = gx * inv;   // d/d world z  (grid x)
  }
}

// L2 prefetch of everything sparse_level will touch for this point: the 8 index entries are loaded (blocking),
// then the (up to) 8 voxel rows are requested with prefetch.global.L2.  Used to pull the next tile's gather
// working set into L2 while the current tile is busy on the tensor core.
__device__ __forceinline__ void sparse_prefetch_l2(const DevScene& sc, int l, float px, float py, float pz) {
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
  int32_t rows[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int xi = (c & 1) ? x1 : x0, yi = (c & 2) ? y1 : y0, zi = (c & 4) ? z1 : z0;
    rows[c] = __ldg(idx + ((size_t)zi * N + yi) * N + xi);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (rows[c] < 0) continue;
    const float4* p = sc.vol8[l] + (size_t)rows[c] * 2;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  }
}

// F.interpolate(mode="bilinear", align_corners=False) source index / weight (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void up_src(int dst, float scale, int in_size, int* i0, int* i1, float* l1) {
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (src < 0.f) src = 0.f;
  int a = (int)src;
  if (a > in_size - 1) a = in_size - 1;
  *i0 = a;
  *i1 = a + ((a < in_size - 1) ? 1 : 0);
  *l1 = __fsub_rn(src, (float)a);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// cross-TU launchers
// ---------------------------------------------------------------------------------------------
struct PointSource {
  // mode 0: explicit points  pts[n][3]
  // mode 1: ray samples via a compact list: id = list[i] (or i if list NULL); p = o[r] + d[r] * mid_z[id]
  // mode 2: tensor-product grid xs[nx] x ys[ny] x zs[nz]
  int mode;
  const float* pts;
  const float* rays_o; const float* rays_d; const float* mid_z; int S;
  const int32_t* list; const int32_t* count;   // device count (NULL -> n)
  const float* xs; const float* ys; const float* zs; int nx, ny, nz;
  int64_t n;
  int sparsify; float fill;
};

// mode: SURF_MLP_* (include/surf_b200.h); the tensor-core editions are used when mode != SURF_MLP_FFMA and n->tc_ok
int launch_sdf_mlp(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_sdf, float* d_grad,
                   bool negate, int mode, cudaStream_t st);
int launch_sdf_tc2(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_sdf, float* d_grad,
                   bool negate, bool fast, cudaStream_t st);
int launch_blend_tc(const surf_net* n, const float* d_feat, const float* d_raydiff, const uint8_t* d_mask, int V,
                    bool packed19, const int32_t* list, const int32_t* count, int64_t n_pts, float* d_rgb,
                    uint8_t* d_views, bool fast, cudaStream_t st);
int launch_sdf_smooth_tc(const surf_scene* s, const surf_net* n, const float* d_pts, int64_t n_pts, const uint8_t* d_flags,
                         float* d_grad, float* d_smooth, bool fast, cudaStream_t st);
bool color_fused_supported(const surf_scene* s, const surf_net* n, int mode);
int launch_color_fused(const surf_scene* s, const surf_net* n, const PointSource& src, float* d_rgb, uint8_t* d_views,
                       bool fast, cudaStream_t st);
int launch_lookup_feature(const surf_scene* s, const PointSource& src, float* d_feat, float* d_raydiff,
                          uint8_t* d_mask, bool packed19, cudaStream_t st, bool small_blocks = false);
int launch_blend(const surf_scene* s_or_null, const surf_net* n, const float* d_feat, const float* d_raydiff,
                 const uint8_t* d_mask_or_null, int V, bool packed19, const int32_t* list, const int32_t* count,
                 int64_t n_pts, float* d_rgb, uint8_t* d_views, int mode, cudaStream_t st);
