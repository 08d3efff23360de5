// Sparse trilinear interpolant with first / mixed-second derivatives along v = (1,1,1): shared by the fp32 second-order
// kernel (sdf_smooth.cu) and its tensor-core edition (sdf_smooth_tc.cu).
#pragma once
#include "surf_internal.cuh"

// one level of the sparse trilinear fetch with its tangent along v = (1,1,1) (world = grid direction, all ones):
// f7 = value, fd7 = J_feat v
__device__ __forceinline__ void sparse_value_tangent(const DevScene& sc, int l, float px, float py, float pz, float* f7,
                                                     float* fd7) {
  const int N = sc.dim[l];
  const float vs = sc.voxel[l];
  const float cx = __fdiv_rn(__fadd_rn(pz, 1.0f), vs), cy = __fdiv_rn(__fadd_rn(py, 1.0f), vs), cz = __fdiv_rn(__fadd_rn(px, 1.0f), vs);
  const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
  const float wx1 = __fsub_rn(cx, fx0), wx0 = __fsub_rn(fx0 + 1.0f, cx);
  const float wy1 = __fsub_rn(cy, fy0), wy0 = __fsub_rn(fy0 + 1.0f, cy);
  const float wz1 = __fsub_rn(cz, fz0), wz0 = __fsub_rn(fz0 + 1.0f, cz);
  const float hi = (float)(N - 1);
  const int x0 = (int)fminf(fmaxf(fx0, 0.f), hi), x1 = (int)fminf(fmaxf(fx0 + 1.0f, 0.f), hi);
  const int y0 = (int)fminf(fmaxf(fy0, 0.f), hi), y1 = (int)fminf(fmaxf(fy0 + 1.0f, 0.f), hi);
  const int z0 = (int)fminf(fmaxf(fz0, 0.f), hi), z1 = (int)fminf(fmaxf(fz0 + 1.0f, 0.f), hi);
  const float inv = 1.0f / vs;
#pragma unroll
  for (int c = 0; c < 7; ++c) { f7[c] = 0.f; fd7[c] = 0.f; }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int xi = (c & 1) ? x1 : x0, yi = (c & 2) ? y1 : y0, zi = (c & 4) ? z1 : z0;
    const int32_t row = __ldg(sc.index[l] + ((size_t)zi * N + yi) * N + xi);
    if (row < 0) continue;
    const float4 a = __ldg(sc.vol8[l] + (size_t)row * 2), b = __ldg(sc.vol8[l] + (size_t)row * 2 + 1);
    const float wx = (c & 1) ? wx1 : wx0, wy = (c & 2) ? wy1 : wy0, wz = (c & 4) ? wz1 : wz0;
    const float sx = (c & 1) ? 1.f : -1.f, sy = (c & 2) ? 1.f : -1.f, sz = (c & 4) ? 1.f : -1.f;
    const float w = wx * wy * wz;
    const float wd = (sx * wy * wz + sy * wx * wz + sz * wx * wy) * inv;
    const float v[7] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z};
#pragma unroll
    for (int k = 0; k < 7; ++k) { f7[k] += v[k] * w; fd7[k] += v[k] * wd; }
  }
}

// reverse of one level: o3 = J_feat^T g (world xyz), od3 = J_feat^T gd + Jd_feat^T g (tangent along (1,1,1))
__device__ __forceinline__ void sparse_back_tangent(const DevScene& sc, int l, float px, float py, float pz, const float* g,
                                                    const float* gd, float* o3, float* od3) {
  const int N = sc.dim[l];
  const float vs = sc.voxel[l];
  const float cx = __fdiv_rn(__fadd_rn(pz, 1.0f), vs), cy = __fdiv_rn(__fadd_rn(py, 1.0f), vs), cz = __fdiv_rn(__fadd_rn(px, 1.0f), vs);
  const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
  const float wx1 = __fsub_rn(cx, fx0), wx0 = __fsub_rn(fx0 + 1.0f, cx);
  const float wy1 = __fsub_rn(cy, fy0), wy0 = __fsub_rn(fy0 + 1.0f, cy);
  const float wz1 = __fsub_rn(cz, fz0), wz0 = __fsub_rn(fz0 + 1.0f, cz);
  const float hi = (float)(N - 1);
  const int x0 = (int)fminf(fmaxf(fx0, 0.f), hi), x1 = (int)fminf(fmaxf(fx0 + 1.0f, 0.f), hi);
  const int y0 = (int)fminf(fmaxf(fy0, 0.f), hi), y1 = (int)fminf(fmaxf(fy0 + 1.0f, 0.f), hi);
  const int z0 = (int)fminf(fmaxf(fz0, 0.f), hi), z1 = (int)fminf(fmaxf(fz0 + 1.0f, 0.f), hi);
  float gx = 0.f, gy = 0.f, gz = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int xi = (c & 1) ? x1 : x0, yi = (c & 2) ? y1 : y0, zi = (c & 4) ? z1 : z0;
    const int32_t row = __ldg(sc.index[l] + ((size_t)zi * N + yi) * N + xi);
    if (row < 0) continue;
    const float4 a = __ldg(sc.vol8[l] + (size_t)row * 2), b = __ldg(sc.vol8[l] + (size_t)row * 2 + 1);
    const float wx = (c & 1) ? wx1 : wx0, wy = (c & 2) ? wy1 : wy0, wz = (c & 4) ? wz1 : wz0;
    const float sx = (c & 1) ? 1.f : -1.f, sy = (c & 2) ? 1.f : -1.f, sz = (c & 4) ? 1.f : -1.f;
    const float s = a.x * g[0] + a.y * g[1] + a.z * g[2] + a.w * g[3] + b.x * g[4] + b.y * g[5] + b.z * g[6];
    const float sd = a.x * gd[0] + a.y * gd[1] + a.z * gd[2] + a.w * gd[3] + b.x * gd[4] + b.y * gd[5] + b.z * gd[6];
    gx += s * (sx * wy * wz); gy += s * (sy * wx * wz); gz += s * (sz * wx * wy);
    // d/d eps of the first derivatives: J^T gd plus the mixed second derivatives (d^2/dxdy = sx sy wz, ...)
    hx += sd * (sx * wy * wz); hy += sd * (sy * wx * wz); hz += sd * (sz * wx * wy);
    hx += s * (sx * (sy * wz + sz * wy)); hy += s * (sy * (sx * wz + sz * wx)); hz += s * (sz * (sx * wy + sy * wx));
  }
  const float inv = 1.0f / vs;
  o3[0] = gz * inv; o3[1] = gy * inv; o3[2] = gx * inv;          // world x <- grid z (projector.py:379)
  // the mixed terms carry 1 / vs^2, the J^T gd terms 1 / vs: split them
  // (recomputed below to keep the two scalings apart)
  od3[0] = hz; od3[1] = hy; od3[2] = hx;
}


// ---------------------------------------------------------------------------------------------
// Batched editions for the tensor-core kernel (one thread owns a point and has registers to spare): the 8 index loads
// of a level are issued together, then the 16 feature loads four corners at a time from clamped rows — 3 dependent
// round trips per level instead of 16.  Same arithmetic in the same order as the functions above.
// ---------------------------------------------------------------------------------------------
struct SparseCorners {
  int N;
  float inv;
  float wx0, wx1, wy0, wy1, wz0, wz1;
  int32_t row[8];
};
__device__ __forceinline__ void sparse_corners(const DevScene& sc, int l, float px, float py, float pz, SparseCorners& C) {
  const int N = sc.dim[l];
  const float vs = sc.voxel[l];
  const float cx = __fdiv_rn(__fadd_rn(pz, 1.0f), vs), cy = __fdiv_rn(__fadd_rn(py, 1.0f), vs), cz = __fdiv_rn(__fadd_rn(px, 1.0f), vs);
  const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
  C.wx1 = __fsub_rn(cx, fx0); C.wx0 = __fsub_rn(fx0 + 1.0f, cx);
  C.wy1 = __fsub_rn(cy, fy0); C.wy0 = __fsub_rn(fy0 + 1.0f, cy);
  C.wz1 = __fsub_rn(cz, fz0); C.wz0 = __fsub_rn(fz0 + 1.0f, cz);
  const float hi = (float)(N - 1);
  const int x0 = (int)fminf(fmaxf(fx0, 0.f), hi), x1 = (int)fminf(fmaxf(fx0 + 1.0f, 0.f), hi);
  const int y0 = (int)fminf(fmaxf(fy0, 0.f), hi), y1 = (int)fminf(fmaxf(fy0 + 1.0f, 0.f), hi);
  const int z0 = (int)fminf(fmaxf(fz0, 0.f), hi), z1 = (int)fminf(fmaxf(fz0 + 1.0f, 0.f), hi);
  C.N = N;
  C.inv = 1.0f / vs;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int xi = (c & 1) ? x1 : x0, yi = (c & 2) ? y1 : y0, zi = (c & 4) ? z1 : z0;
    C.row[c] = __ldg(sc.index[l] + ((size_t)zi * N + yi) * N + xi);
  }
}

// f7 = value, fd7 = J_feat v  (sparse_value_tangent)
__device__ __forceinline__ void sparse_value_tangent_batched(const DevScene& sc, int l, float px, float py, float pz,
                                                             float* f7, float* fd7) {
  SparseCorners C;
  sparse_corners(sc, l, px, py, pz, C);
#pragma unroll
  for (int c = 0; c < 7; ++c) { f7[c] = 0.f; fd7[c] = 0.f; }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4 a[4], b[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int32_t row = C.row[h * 4 + cc] < 0 ? 0 : C.row[h * 4 + cc];
      a[cc] = __ldg(sc.vol8[l] + (size_t)row * 2);
      b[cc] = __ldg(sc.vol8[l] + (size_t)row * 2 + 1);
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = h * 4 + cc;
      if (C.row[c] < 0) continue;
      const float wx = (c & 1) ? C.wx1 : C.wx0, wy = (c & 2) ? C.wy1 : C.wy0, wz = (c & 4) ? C.wz1 : C.wz0;
      const float sx = (c & 1) ? 1.f : -1.f, sy = (c & 2) ? 1.f : -1.f, sz = (c & 4) ? 1.f : -1.f;
      const float w = wx * wy * wz;
      const float wd = (sx * wy * wz + sy * wx * wz + sz * wx * wy) * C.inv;
      const float v[7] = {a[cc].x, a[cc].y, a[cc].z, a[cc].w, b[cc].x, b[cc].y, b[cc].z};
#pragma unroll
      for (int k = 0; k < 7; ++k) { f7[k] += v[k] * w; fd7[k] += v[k] * wd; }
    }
  }
}

// Both reverse passes of one level in one gather: g3 = J^T g / vs, gd3 = J^T gd / vs (world xyz), m3 = the mixed
// second derivatives contracted with g, unscaled (the caller multiplies by 1 / vs^2).  Equals
// sparse_back_tangent(g, 0) -> (g3, m3) and sparse_back_tangent(gd, 0) -> (gd3, .).
__device__ __forceinline__ void sparse_back_fused(const DevScene& sc, int l, float px, float py, float pz, const float* g,
                                                  const float* gd, float* g3, float* gd3, float* m3) {
  SparseCorners C;
  sparse_corners(sc, l, px, py, pz, C);
  float gx = 0.f, gy = 0.f, gz = 0.f, hx = 0.f, hy = 0.f, hz = 0.f, mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4 a[4], b[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int32_t row = C.row[h * 4 + cc] < 0 ? 0 : C.row[h * 4 + cc];
      a[cc] = __ldg(sc.vol8[l] + (size_t)row * 2);
      b[cc] = __ldg(sc.vol8[l] + (size_t)row * 2 + 1);
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = h * 4 + cc;
      if (C.row[c] < 0) continue;
      const float wx = (c & 1) ? C.wx1 : C.wx0, wy = (c & 2) ? C.wy1 : C.wy0, wz = (c & 4) ? C.wz1 : C.wz0;
      const float sx = (c & 1) ? 1.f : -1.f, sy = (c & 2) ? 1.f : -1.f, sz = (c & 4) ? 1.f : -1.f;
      const float4 A = a[cc], B = b[cc];
      const float s = A.x * g[0] + A.y * g[1] + A.z * g[2] + A.w * g[3] + B.x * g[4] + B.y * g[5] + B.z * g[6];
      const float sd = A.x * gd[0] + A.y * gd[1] + A.z * gd[2] + A.w * gd[3] + B.x * gd[4] + B.y * gd[5] + B.z * gd[6];
      gx += s * (sx * wy * wz); gy += s * (sy * wx * wz); gz += s * (sz * wx * wy);
      hx += sd * (sx * wy * wz); hy += sd * (sy * wx * wz); hz += sd * (sz * wx * wy);
      mx += s * (sx * (sy * wz + sz * wy)); my += s * (sy * (sx * wz + sz * wx)); mz += s * (sz * (sx * wy + sy * wx));
    }
  }
  g3[0] = gz * C.inv; g3[1] = gy * C.inv; g3[2] = gx * C.inv;          // world x <- grid z (projector.py:379)
  gd3[0] = hz * C.inv; gd3[1] = hy * C.inv; gd3[2] = hx * C.inv;
  m3[0] = mz; m3[1] = my; m3[2] = mx;
}

// first-order reverse pass of one level, batched loads: g3 = J_feat^T g / vs (world xyz)
__device__ __forceinline__ void sparse_back_first(const DevScene& sc, int l, float px, float py, float pz, const float* g,
                                                  float* g3) {
  SparseCorners C;
  sparse_corners(sc, l, px, py, pz, C);
  float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4 a[4], b[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int32_t row = C.row[h * 4 + cc] < 0 ? 0 : C.row[h * 4 + cc];
      a[cc] = __ldg(sc.vol8[l] + (size_t)row * 2);
      b[cc] = __ldg(sc.vol8[l] + (size_t)row * 2 + 1);
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = h * 4 + cc;
      if (C.row[c] < 0) continue;
      const float wx = (c & 1) ? C.wx1 : C.wx0, wy = (c & 2) ? C.wy1 : C.wy0, wz = (c & 4) ? C.wz1 : C.wz0;
      const float sx = (c & 1) ? 1.f : -1.f, sy = (c & 2) ? 1.f : -1.f, sz = (c & 4) ? 1.f : -1.f;
      const float4 A = a[cc], B = b[cc];
      const float s = A.x * g[0] + A.y * g[1] + A.z * g[2] + A.w * g[3] + B.x * g[4] + B.y * g[5] + B.z * g[6];
      gx += s * (sx * wy * wz); gy += s * (sy * wx * wz); gz += s * (sz * wx * wy);
    }
  }
  g3[0] = gz * C.inv; g3[1] = gy * C.inv; g3[2] = gx * C.inv;          // world x <- grid z (projector.py:379)
}

// value only (forward), batched loads
__device__ __forceinline__ void sparse_value_batched(const DevScene& sc, int l, float px, float py, float pz, float* f7) {
  SparseCorners C;
  sparse_corners(sc, l, px, py, pz, C);
#pragma unroll
  for (int c = 0; c < 7; ++c) f7[c] = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4 a[4], b[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int32_t row = C.row[h * 4 + cc] < 0 ? 0 : C.row[h * 4 + cc];
      a[cc] = __ldg(sc.vol8[l] + (size_t)row * 2);
      b[cc] = __ldg(sc.vol8[l] + (size_t)row * 2 + 1);
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = h * 4 + cc;
      if (C.row[c] < 0) continue;
      const float wx = (c & 1) ? C.wx1 : C.wx0, wy = (c & 2) ? C.wy1 : C.wy0, wz = (c & 4) ? C.wz1 : C.wz0;
      const float w = wx * wy * wz;
      const float v[7] = {a[cc].x, a[cc].y, a[cc].z, a[cc].w, b[cc].x, b[cc].y, b[cc].z};
#pragma unroll
      for (int k = 0; k < 7; ++k) f7[k] += v[k] * w;
    }
  }
}
