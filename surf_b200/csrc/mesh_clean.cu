// Mesh cleaning after extract_geometry (utils/clean_mesh.py:10-129, called from runner.py:233 with --clean_mesh):
//   1. the object masks are dilated with a disk (skimage binary_dilation, :119-123)              k_mask_row_prefix + k_mask_dilate
//   2. faces whose three vertices are seen inside the dilated masks of more than `min_nb_visible` views stay
//      (clean_mesh_by_mask, :10-34)                                                              k_vertex_visibility
//   3. faces that are the FIRST hit of no masked camera ray go (clean_mesh_outside_frustum, :38-96; the reference casts
//      h*up x w*up rays per view through trimesh / embree).  Camera rays through a regular sample grid are a z-buffer:
//      every face is projected, the samples inside its screen bounding box get an exact ray / triangle test and the
//      nearest hit per sample wins by a 64-bit atomicMin on (t, face)                            k_raster_faces + k_raster_collect
//   4. connected components of the face-adjacency graph with fewer than `min_len` faces go (:99-104; an edge shared by
//      exactly two faces links them, trimesh.graph.face_adjacency): edge hash table + lock-free union-find
//                                                                                                k_edge_insert, k_edge_union, k_cc_*
// Everything is integer / byte work or a 3x3 projection per element: HBM / atomic bound, no tensor cores.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "surf_internal.cuh"

static inline int mc_grid(int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  const int64_t cap = (int64_t)surf_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------
// 1. binary dilation with a disk of radius R (skimage.morphology.disk: dx^2 + dy^2 <= R^2), zero border
// ---------------------------------------------------------------------------------------------
// inclusive row prefix sums of the binary masks: P[v][y][x] = #set pixels in row y up to x (one warp per row)
__global__ void k_mask_row_prefix(const uint8_t* __restrict__ mask, int rows, int w, int32_t* __restrict__ pre) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += gridDim.x * (blockDim.x >> 5)) {
    const uint8_t* m = mask + (size_t)row * w;
    int32_t* p = pre + (size_t)row * w;
    int carry = 0;
    for (int x0 = 0; x0 < w; x0 += 32) {
      const int x = x0 + lane;
      int v = (x < w && m[x]) ? 1 : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (x < w) p[x] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
}
// out = 1 iff some row y + dy has a set pixel within the half width of the disk at dy
__global__ void k_mask_dilate(const int32_t* __restrict__ pre, int nv, int h, int w, int R, uint8_t* __restrict__ out) {
  const int64_t n = (int64_t)nv * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const int64_t img = i / ((int64_t)w * h);
    int hit = 0;
    for (int dy = -R; dy <= R && !hit; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      int half = (int)floorf(sqrtf((float)(R * R - dy * dy)));
      while ((half + 1) * (half + 1) + dy * dy <= R * R) ++half;           // guard the sqrt rounding
      while (half * half + dy * dy > R * R) --half;
      const int xa = x - half - 1, xb = min(x + half, w - 1);
      const int32_t* p = pre + (img * h + yy) * (int64_t)w;
      const int cnt = p[xb] - (xa >= 0 ? p[xa] : 0);
      hit = cnt > 0;
    }
    out[i] = hit ? 1 : 0;
  }
}

extern "C" int surf_mask_dilate(const uint8_t* d_masks, int32_t n_views, int32_t h, int32_t w, int32_t radius,
                                int32_t* d_workspace, uint8_t* d_out, void* stream) {
  SURF_CHECK_ARG(d_masks && d_workspace && d_out, "null pointer");
  SURF_CHECK_ARG(n_views >= 1 && h >= 1 && w >= 1 && radius >= 0 && radius <= 4096, "mask shape / radius");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = n_views * h;
  k_mask_row_prefix<<<mc_grid(rows, 8), 256, 0, st>>>(d_masks, rows, w, d_workspace);
  SURF_LAUNCH_CHECK();
  k_mask_dilate<<<mc_grid((int64_t)rows * w, 256), 256, 0, st>>>(d_workspace, n_views, h, w, radius, d_out);
  SURF_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// 2. per vertex: in how many views does it project inside the image AND onto the (dilated) mask
//    (pts_cam = inv(c2w) p, pts_img = K pts_cam, bilinear tap of the mask with align_corners=True, zero padding)
// ---------------------------------------------------------------------------------------------
struct MeshViews {
  float w2c[SURF_MAX_VIEWS + 1][12];
  float K[SURF_MAX_VIEWS + 1][9];
  int n;
};

__global__ void k_vertex_visibility(const float* __restrict__ verts, int64_t nvert, const MeshViews V,
                                    const uint8_t* __restrict__ masks, int h, int w, int32_t* __restrict__ count,
                                    int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvert; i += (int64_t)gridDim.x * blockDim.x) {
    const float px = verts[i * 3], py = verts[i * 3 + 1], pz = verts[i * 3 + 2];
    int c = 0;
    for (int v = 0; v < V.n; ++v) {
      const float* M = V.w2c[v];
      const float cx = M[0] * px + M[1] * py + M[2] * pz + M[3];
      const float cy = M[4] * px + M[5] * py + M[6] * pz + M[7];
      const float cz = M[8] * px + M[9] * py + M[10] * pz + M[11];
      const float* K = V.K[v];
      const float ix = K[0] * cx + K[1] * cy + K[2] * cz;
      const float iy = K[3] * cx + K[4] * cy + K[5] * cz;
      const float iz = K[6] * cx + K[7] * cy + K[8] * cz;
      const float zc = fmaxf(iz, 1e-8f);
      float gx = 2.0f * (ix / zc) / (float)(w - 1) - 1.0f;
      float gy = 2.0f * (iy / zc) / (float)(h - 1) - 1.0f;
      const bool in_img = fabsf(gx) <= 1.0f && fabsf(gy) <= 1.0f && iz > 1e-8f;
      gx = fminf(fmaxf(gx, -10.f), 10.f);
      gy = fminf(fmaxf(gy, -10.f), 10.f);
      // grid_sample, bilinear, align_corners=True, padding zeros
      const float sx = ((gx + 1.0f) * 0.5f) * (float)(w - 1), sy = ((gy + 1.0f) * 0.5f) * (float)(h - 1);
      const float fx = floorf(sx), fy = floorf(sy);
      const int x0 = (int)fx, y0 = (int)fy;
      const float tx = sx - fx, ty = sy - fy;
      const uint8_t* m = masks + (size_t)v * h * w;
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int xi = x0 + (t & 1), yi = y0 + (t >> 1);
        if (xi < 0 || xi >= w || yi < 0 || yi >= h) continue;
        const float wgt = ((t & 1) ? tx : 1.0f - tx) * ((t & 2) ? ty : 1.0f - ty);
        acc += (m[(size_t)yi * w + xi] ? 1.0f : 0.0f) * wgt;
      }
      c += (in_img && acc > 0.f) ? 1 : 0;
    }
    count[i] = accumulate ? count[i] + c : c;
  }
}

extern "C" int surf_mesh_vertex_visibility(const float* d_vertices, int64_t n_vertices, const float* h_w2c,
                                           const float* h_K, int32_t n_views, const uint8_t* d_masks, int32_t h,
                                           int32_t w, int32_t* d_count, void* stream) {
  if (n_vertices <= 0) return 0;
  SURF_CHECK_ARG(d_vertices && h_w2c && h_K && d_masks && d_count, "null pointer");
  SURF_CHECK_ARG(n_views >= 1, "n_views");
  // the view parameters travel as a kernel argument: SURF_MAX_VIEWS + 1 views per launch, counts accumulated
  for (int v0 = 0; v0 < n_views; v0 += SURF_MAX_VIEWS + 1) {
    MeshViews V;
    memset(&V, 0, sizeof(V));
    V.n = n_views - v0 < SURF_MAX_VIEWS + 1 ? n_views - v0 : SURF_MAX_VIEWS + 1;
    for (int v = 0; v < V.n; ++v) {
      memcpy(V.w2c[v], h_w2c + (size_t)(v0 + v) * 12, 12 * sizeof(float));
      memcpy(V.K[v], h_K + (size_t)(v0 + v) * 9, 9 * sizeof(float));
    }
    k_vertex_visibility<<<mc_grid(n_vertices, 256), 256, 0, (cudaStream_t)stream>>>(
        d_vertices, n_vertices, V, d_masks + (size_t)v0 * h * w, h, w, d_count, v0 > 0 ? 1 : 0);
    SURF_LAUNCH_CHECK();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// 3. first hit of the camera rays of one view (rays through the hs x ws sample grid, linspace(0, h-1, hs) etc.)
// ---------------------------------------------------------------------------------------------
struct RasterView {
  float w2c[12];     // world -> camera
  float c2w[12];     // camera -> world
  float K[9], Kinv[9];
  int h, w, hs, ws;
};
#define RASTER_EMPTY 0xffffffffffffffffull
#define RASTER_BIG 4096          // faces covering more samples go to the block-per-face kernel

// world-space ray through sample (i, j) exactly as the reference builds it (clean_mesh.py:50-64)
__device__ __forceinline__ void raster_ray(const RasterView& V, int i, int j, float* d) {
  const float x = V.ws > 1 ? (float)j * ((float)(V.w - 1) / (float)(V.ws - 1)) : 0.f;
  const float y = V.hs > 1 ? (float)i * ((float)(V.h - 1) / (float)(V.hs - 1)) : 0.f;
  const float* Ki = V.Kinv;
  float px = Ki[0] * x + Ki[1] * y + Ki[2], py = Ki[3] * x + Ki[4] * y + Ki[5], pz = Ki[6] * x + Ki[7] * y + Ki[8];
  const float inv = 1.0f / sqrtf(px * px + py * py + pz * pz);
  px *= inv; py *= inv; pz *= inv;
  const float* M = V.c2w;
  d[0] = M[0] * px + M[1] * py + M[2] * pz;
  d[1] = M[4] * px + M[5] * py + M[6] * pz;
  d[2] = M[8] * px + M[9] * py + M[10] * pz;
}

// Moller-Trumbore, both sides, t > 0
__device__ __forceinline__ bool ray_tri(const float* o, const float* d, const float* a, const float* e1, const float* e2,
                                        float& t) {
  const float px = d[1] * e2[2] - d[2] * e2[1], py = d[2] * e2[0] - d[0] * e2[2], pz = d[0] * e2[1] - d[1] * e2[0];
  const float det = e1[0] * px + e1[1] * py + e1[2] * pz;
  if (det == 0.f) return false;
  const float inv = 1.0f / det;
  const float tx = o[0] - a[0], ty = o[1] - a[1], tz = o[2] - a[2];
  const float u = (tx * px + ty * py + tz * pz) * inv;
  if (u < 0.f || u > 1.f) return false;
  const float qx = ty * e1[2] - tz * e1[1], qy = tz * e1[0] - tx * e1[2], qz = tx * e1[1] - ty * e1[0];
  const float v = (d[0] * qx + d[1] * qy + d[2] * qz) * inv;
  if (v < 0.f || u + v > 1.f) return false;
  t = (e2[0] * qx + e2[1] * qy + e2[2] * qz) * inv;
  return t > 0.f;
}

struct FaceSetup {
  float a[3], e1[3], e2[3];
  int i0, i1, j0, j1;     // sample bounding box (inclusive), empty when i0 > i1
};
__device__ __forceinline__ void face_setup(const RasterView& V, const float* __restrict__ verts,
                                           const int32_t* __restrict__ faces, int64_t f, FaceSetup& S) {
  float p[3][3];
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  bool behind = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int64_t vi = faces[f * 3 + k];
    p[k][0] = verts[vi * 3]; p[k][1] = verts[vi * 3 + 1]; p[k][2] = verts[vi * 3 + 2];
    const float* M = V.w2c;
    const float cx = M[0] * p[k][0] + M[1] * p[k][1] + M[2] * p[k][2] + M[3];
    const float cy = M[4] * p[k][0] + M[5] * p[k][1] + M[6] * p[k][2] + M[7];
    const float cz = M[8] * p[k][0] + M[9] * p[k][1] + M[10] * p[k][2] + M[11];
    const float ix = V.K[0] * cx + V.K[1] * cy + V.K[2] * cz, iy = V.K[3] * cx + V.K[4] * cy + V.K[5] * cz;
    const float iz = V.K[6] * cx + V.K[7] * cy + V.K[8] * cz;
    if (!(iz > 1e-6f)) behind = true;
    const float x = ix / iz, y = iy / iz;
    xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { S.a[k] = p[0][k]; S.e1[k] = p[1][k] - p[0][k]; S.e2[k] = p[2][k] - p[0][k]; }
  if (behind) {      // a vertex at or behind the camera plane: test every sample
    S.i0 = 0; S.i1 = V.hs - 1; S.j0 = 0; S.j1 = V.ws - 1;
    return;
  }
  const float sx = V.w > 1 ? (float)(V.ws - 1) / (float)(V.w - 1) : 0.f, sy = V.h > 1 ? (float)(V.hs - 1) / (float)(V.h - 1) : 0.f;
  // one sample of slack on each side covers the rounding of the projection
  const float j0 = floorf(xmin * sx) - 1.f, j1 = ceilf(xmax * sx) + 1.f, i0 = floorf(ymin * sy) - 1.f, i1 = ceilf(ymax * sy) + 1.f;
  if (!(j1 >= 0.f) || !(i1 >= 0.f) || !(j0 <= (float)(V.ws - 1)) || !(i0 <= (float)(V.hs - 1))) {
    S.i0 = 1; S.i1 = 0; S.j0 = 1; S.j1 = 0;      // off screen (or NaN)
    return;
  }
  S.j0 = (int)fmaxf(j0, 0.f); S.j1 = (int)fminf(j1, (float)(V.ws - 1));
  S.i0 = (int)fmaxf(i0, 0.f); S.i1 = (int)fminf(i1, (float)(V.hs - 1));
}
__device__ __forceinline__ void raster_sample(const RasterView& V, const FaceSetup& S, const uint8_t* __restrict__ mask,
                                              int i, int j, int64_t f, unsigned long long* __restrict__ zbuf) {
  // the ray of a sample counts only where the nearest-upsampled mask is set (F.interpolate nearest: src = floor(dst * in / out))
  const int mi = min((int)(((int64_t)i * V.h) / V.hs), V.h - 1), mj = min((int)(((int64_t)j * V.w) / V.ws), V.w - 1);
  if (!mask[(size_t)mi * V.w + mj]) return;
  float d[3];
  raster_ray(V, i, j, d);
  const float o[3] = {V.c2w[3], V.c2w[7], V.c2w[11]};
  float t;
  if (!ray_tri(o, d, S.a, S.e1, S.e2, t)) return;
  const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)(uint32_t)f;
  atomicMin(zbuf + (size_t)i * V.ws + j, key);
}

__global__ void k_raster_faces(const RasterView V, const float* __restrict__ verts, const int32_t* __restrict__ faces,
                               int64_t nf, const uint8_t* __restrict__ mask, unsigned long long* __restrict__ zbuf,
                               int32_t* __restrict__ big_list, int32_t* __restrict__ big_count, int big_cap) {
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    FaceSetup S;
    face_setup(V, verts, faces, f, S);
    if (S.i0 > S.i1) continue;
    const int64_t area = (int64_t)(S.i1 - S.i0 + 1) * (S.j1 - S.j0 + 1);
    if (area > RASTER_BIG) {
      const int slot = atomicAdd(big_count, 1);
      if (slot < big_cap) big_list[slot] = (int32_t)f;
      continue;
    }
    for (int i = S.i0; i <= S.i1; ++i)
      for (int j = S.j0; j <= S.j1; ++j) raster_sample(V, S, mask, i, j, f, zbuf);
  }
}
// the few faces with a large footprint: one block per face, threads over its samples
__global__ void k_raster_big(const RasterView V, const float* __restrict__ verts, const int32_t* __restrict__ faces,
                             const uint8_t* __restrict__ mask, unsigned long long* __restrict__ zbuf,
                             const int32_t* __restrict__ big_list, const int32_t* __restrict__ big_count, int big_cap) {
  const int n = min(*big_count, big_cap);
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    const int64_t f = big_list[b];
    FaceSetup S;
    face_setup(V, verts, faces, f, S);
    if (S.i0 > S.i1) continue;
    const int wj = S.j1 - S.j0 + 1;
    const int64_t area = (int64_t)(S.i1 - S.i0 + 1) * wj;
    for (int64_t s = threadIdx.x; s < area; s += blockDim.x)
      raster_sample(V, S, mask, S.i0 + (int)(s / wj), S.j0 + (int)(s % wj), f, zbuf);
  }
}
__global__ void k_fill_u64(unsigned long long* __restrict__ p, int64_t n, unsigned long long v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
// face_hit[f] = 1 for every face that is the first hit of a masked sample; stats[0] += masked samples without a hit
__global__ void k_raster_collect(const RasterView V, const uint8_t* __restrict__ mask,
                                 const unsigned long long* __restrict__ zbuf, uint8_t* __restrict__ face_hit,
                                 int32_t* __restrict__ stats) {
  const int64_t n = (int64_t)V.hs * V.ws;
  int miss = 0;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(s / V.ws), j = (int)(s % V.ws);
    const int mi = min((int)(((int64_t)i * V.h) / V.hs), V.h - 1), mj = min((int)(((int64_t)j * V.w) / V.ws), V.w - 1);
    if (!mask[(size_t)mi * V.w + mj]) continue;
    const unsigned long long k = zbuf[s];
    if (k == RASTER_EMPTY) ++miss;
    else face_hit[(uint32_t)(k & 0xffffffffull)] = 1;
  }
  if (miss) atomicAdd(stats, miss);
}

static void invert3(const float* m, float* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  const double r = 1.0 / det;
  o[0] = (float)((e * i - f * h) * r); o[1] = (float)((c * h - b * i) * r); o[2] = (float)((b * f - c * e) * r);
  o[3] = (float)((f * g - d * i) * r); o[4] = (float)((a * i - c * g) * r); o[5] = (float)((c * d - a * f) * r);
  o[6] = (float)((d * h - e * g) * r); o[7] = (float)((b * g - a * h) * r); o[8] = (float)((a * e - b * d) * r);
}

static size_t raster_workspace_bytes(int32_t hs, int32_t ws) { return (size_t)hs * ws * 8 + (size_t)(65536 + 4) * 4; }
extern "C" size_t surf_mesh_raster_workspace_bytes(int32_t hs, int32_t ws) { return raster_workspace_bytes(hs, ws); }

extern "C" int surf_mesh_first_hits(const float* d_vertices, const int32_t* d_faces, int64_t n_faces, const float* h_w2c,
                                    const float* h_c2w, const float* h_K, const uint8_t* d_mask, int32_t h, int32_t w,
                                    int32_t hs, int32_t ws, void* d_workspace, size_t workspace_bytes,
                                    uint8_t* d_face_hit, int32_t* d_stats, void* stream) {
  SURF_CHECK_ARG(d_vertices && d_faces && h_w2c && h_c2w && h_K && d_mask && d_workspace && d_face_hit && d_stats, "null pointer");
  SURF_CHECK_ARG(h >= 1 && w >= 1 && hs >= 1 && ws >= 1, "image shape");
  SURF_CHECK_ARG(n_faces < 0x7fffffffll, "too many faces");
  SURF_CHECK_ARG(workspace_bytes >= raster_workspace_bytes(hs, ws), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  RasterView V;
  memcpy(V.w2c, h_w2c, sizeof(V.w2c));
  memcpy(V.c2w, h_c2w, sizeof(V.c2w));
  memcpy(V.K, h_K, sizeof(V.K));
  invert3(h_K, V.Kinv);
  V.h = h; V.w = w; V.hs = hs; V.ws = ws;
  unsigned long long* zbuf = (unsigned long long*)d_workspace;
  int32_t* big_list = (int32_t*)((uint8_t*)d_workspace + (size_t)hs * ws * 8);
  int32_t* big_count = big_list + 65536;
  const int64_t ns = (int64_t)hs * ws;
  k_fill_u64<<<mc_grid(ns, 256), 256, 0, st>>>(zbuf, ns, RASTER_EMPTY);
  SURF_LAUNCH_CHECK();
  SURF_CUDA(cudaMemsetAsync(big_count, 0, 4 * sizeof(int32_t), st));
  if (n_faces > 0) {
    k_raster_faces<<<mc_grid(n_faces, 128), 128, 0, st>>>(V, d_vertices, d_faces, n_faces, d_mask, zbuf, big_list, big_count, 65536);
    SURF_LAUNCH_CHECK();
    k_raster_big<<<surf_num_sms() * 4, 256, 0, st>>>(V, d_vertices, d_faces, d_mask, zbuf, big_list, big_count, 65536);
    SURF_LAUNCH_CHECK();
  }
  k_raster_collect<<<mc_grid(ns, 256), 256, 0, st>>>(V, d_mask, zbuf, d_face_hit, d_stats);
  SURF_LAUNCH_CHECK();
  // more than 65536 large faces cannot be rasterised completely: report through stats[1]
  SURF_CUDA(cudaMemcpyAsync(d_stats + 1, big_count, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// 4. connected components of the face adjacency graph
// ---------------------------------------------------------------------------------------------
#define EDGE_EMPTY 0xffffffffffffffffull
struct EdgeTable {
  unsigned long long* key;   // [cap]
  int32_t* cnt;              // [cap]
  int32_t* f0;               // [cap]
  int32_t* f1;               // [cap]
  uint64_t mask;             // cap - 1
};
__device__ __forceinline__ uint64_t edge_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}
__global__ void k_edge_insert(const int32_t* __restrict__ faces, int64_t nf, EdgeTable T) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nf * 3; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = e / 3;
    const int k = (int)(e - f * 3);
    const uint32_t a = (uint32_t)faces[f * 3 + k], b = (uint32_t)faces[f * 3 + (k + 1) % 3];
    if (a == b) continue;          // degenerate edge
    const unsigned long long key = ((unsigned long long)min(a, b) << 32) | max(a, b);
    uint64_t slot = edge_hash(key) & T.mask;
    while (true) {
      const unsigned long long old = atomicCAS(T.key + slot, EDGE_EMPTY, key);
      if (old == EDGE_EMPTY || old == key) {
        const int n = atomicAdd(T.cnt + slot, 1);
        if (n == 0) T.f0[slot] = (int32_t)f;
        else if (n == 1) T.f1[slot] = (int32_t)f;
        break;
      }
      slot = (slot + 1) & T.mask;
    }
  }
}
__device__ __forceinline__ int32_t uf_find(int32_t* parent, int32_t x) {
  while (true) {
    const int32_t p = ((volatile int32_t*)parent)[x];
    if (p == x) return x;
    x = p;
  }
}
__global__ void k_uf_init(int32_t* __restrict__ parent, int32_t* __restrict__ linked, int64_t nf) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (int64_t)gridDim.x * blockDim.x) {
    parent[i] = (int32_t)i;
    linked[i] = 0;
  }
}
// an edge shared by exactly two faces links them (trimesh face_adjacency); roots are hooked larger -> smaller
__global__ void k_edge_union(EdgeTable T, int32_t* __restrict__ parent, int32_t* __restrict__ linked) {
  const int64_t cap = (int64_t)T.mask + 1;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += (int64_t)gridDim.x * blockDim.x) {
    if (T.cnt[s] != 2) continue;
    int32_t a = T.f0[s], b = T.f1[s];
    linked[a] = 1;
    linked[b] = 1;
    while (true) {
      a = uf_find(parent, a);
      b = uf_find(parent, b);
      if (a == b) break;
      const int32_t hi = a > b ? a : b, lo = a > b ? b : a;
      if (atomicCAS(parent + hi, hi, lo) == hi) break;
    }
  }
}
__global__ void k_cc_label(int32_t* __restrict__ parent, int32_t* __restrict__ label, int32_t* __restrict__ size, int64_t nf) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = uf_find(parent, (int32_t)i);
    label[i] = r;
    atomicAdd(size + r, 1);
  }
}
__global__ void k_cc_keep(const int32_t* __restrict__ label, const int32_t* __restrict__ size,
                          const int32_t* __restrict__ linked, int64_t nf, int min_len, uint8_t* __restrict__ keep) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (int64_t)gridDim.x * blockDim.x)
    keep[i] = (linked[i] && size[label[i]] >= min_len) ? 1 : 0;      // faces without a neighbour are not graph nodes
}

static size_t cc_table_cap(int64_t nf) {
  size_t cap = 1024;
  while (cap < (size_t)nf * 3 * 2) cap <<= 1;
  return cap;
}
extern "C" size_t surf_mesh_components_workspace_bytes(int64_t n_faces) {
  const size_t cap = cc_table_cap(n_faces < 1 ? 1 : n_faces);
  return cap * (8 + 4 + 4 + 4) + (size_t)(n_faces < 1 ? 1 : n_faces) * 4 * 3 + 256;
}
extern "C" int surf_mesh_components(const int32_t* d_faces, int64_t n_faces, int32_t min_len, void* d_workspace,
                                    size_t workspace_bytes, int32_t* d_label, uint8_t* d_keep, void* stream) {
  if (n_faces <= 0) return 0;
  SURF_CHECK_ARG(d_faces && d_workspace && d_label && d_keep, "null pointer");
  SURF_CHECK_ARG(n_faces < 0x7fffffffll / 3, "too many faces");
  SURF_CHECK_ARG(workspace_bytes >= surf_mesh_components_workspace_bytes(n_faces), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cap = cc_table_cap(n_faces);
  uint8_t* p = (uint8_t*)d_workspace;
  EdgeTable T;
  T.key = (unsigned long long*)p; p += cap * 8;
  T.cnt = (int32_t*)p; p += cap * 4;
  T.f0 = (int32_t*)p; p += cap * 4;
  T.f1 = (int32_t*)p; p += cap * 4;
  T.mask = cap - 1;
  int32_t* parent = (int32_t*)p; p += (size_t)n_faces * 4;
  int32_t* size = (int32_t*)p; p += (size_t)n_faces * 4;
  int32_t* linked = (int32_t*)p;
  SURF_CUDA(cudaMemsetAsync(T.key, 0xff, cap * 8, st));
  SURF_CUDA(cudaMemsetAsync(T.cnt, 0, cap * 4, st));
  SURF_CUDA(cudaMemsetAsync(size, 0, (size_t)n_faces * 4, st));
  k_uf_init<<<mc_grid(n_faces, 256), 256, 0, st>>>(parent, linked, n_faces);
  SURF_LAUNCH_CHECK();
  k_edge_insert<<<mc_grid(n_faces * 3, 256), 256, 0, st>>>(d_faces, n_faces, T);
  SURF_LAUNCH_CHECK();
  k_edge_union<<<mc_grid((int64_t)cap, 256), 256, 0, st>>>(T, parent, linked);
  SURF_LAUNCH_CHECK();
  k_cc_label<<<mc_grid(n_faces, 256), 256, 0, st>>>(parent, d_label, size, n_faces);
  SURF_LAUNCH_CHECK();
  k_cc_keep<<<mc_grid(n_faces, 256), 256, 0, st>>>(d_label, size, linked, n_faces, min_len, d_keep);
  SURF_LAUNCH_CHECK();
  return 0;
}
