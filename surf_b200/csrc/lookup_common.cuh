// Projection of a sample point into ONE source view + bilinear gather of RGB / 4 pyramid levels + compute_angle
// (lookup_feature + compute_angle, projector.py:485-556, quirk Q14).  Shared by the stand-alone gather kernel
// (blend.cu: k_lookup_feature) and the producer warps of the fused colour kernel (blend_tc.cu: k_color_fused).
#pragma once
#include "surf_internal.cuh"

#define FEAT_REC 20   // internal record per (point, view): 19 channels + validity flag

__device__ __forceinline__ void bilinear_taps(float x, float y, int w, int h, int& x0, int& y0, float& tx, float& ty,
                                              bool& ok) {
  // grid = x / ((w-1)/2) - 1 (projector.py:533), sampled with align_corners=False (:544)
  const float gx = x / ((float)(w - 1) * 0.5f) - 1.0f;
  const float gy = y / ((float)(h - 1) * 0.5f) - 1.0f;
  const float ix = (gx + 1.0f) * ((float)w * 0.5f) - 0.5f;
  const float iy = (gy + 1.0f) * ((float)h * 0.5f) - 0.5f;
  const float fx = floorf(ix), fy = floorf(iy);
  ok = (fx >= -1.f) && (fx < (float)w) && (fy >= -1.f) && (fy < (float)h);   // false for NaN / far away
  x0 = ok ? (int)fx : 0;
  y0 = ok ? (int)fy : 0;
  tx = ix - fx;
  ty = iy - fy;
}

// rec[0..18] = [rgb3, f0(4), f1(4), f2(4), f3(4)], rec[19] = validity (1 / 0); rd = (normalised direction difference, dot)
__device__ __forceinline__ void lookup_row(const DevScene& sc, float px, float py, float pz, int v, float (&rec)[FEAT_REC],
                                           float4& rd) {
  // ---- compute_angle (projector.py:485-498) ----
  float ax = sc.refcen[0] - px, ay = sc.refcen[1] - py, az = sc.refcen[2] - pz;
  float inv = 1.0f / (sqrtf(ax * ax + ay * ay + az * az) + 1e-6f);
  ax *= inv; ay *= inv; az *= inv;
  float bx = sc.cen[v][0] - px, by = sc.cen[v][1] - py, bz = sc.cen[v][2] - pz;
  inv = 1.0f / (sqrtf(bx * bx + by * by + bz * bz) + 1e-6f);
  bx *= inv; by *= inv; bz *= inv;
  const float ddx = ax - bx, ddy = ay - by, ddz = az - bz;
  const float dn = fmaxf(sqrtf(ddx * ddx + ddy * ddy + ddz * ddz), 1e-6f);
  const float dot = ax * bx + ay * by + az * bz;
  // ---- projection ----
  const float* M = sc.w2c[v];
  const float cx = M[0] * px + M[1] * py + M[2] * pz + M[3];
  const float cy = M[4] * px + M[5] * py + M[6] * pz + M[7];
  const float cz = M[8] * px + M[9] * py + M[10] * pz + M[11];
  const float* K = sc.K[v];
  const float w = K[6] * cx + K[7] * cy + K[8] * cz;
  bool valid = w > 0.f;
  float scale = 1.0f;
#pragma unroll
  for (int lv = 0; lv < 4; ++lv) {
    const int fw = sc.fw[lv], fh = sc.fh[lv];
    const float u = (K[0] * scale) * cx + (K[1] * scale) * cy + (K[2] * scale) * cz;
    const float vv = (K[3] * scale) * cx + (K[4] * scale) * cy + (K[5] * scale) * cz;
    const float x = u / w, y = vv / w;
    valid = valid && (x >= 0.f) && (x < (float)fw) && (y >= 0.f) && (y < (float)fh);
    int x0, y0;
    float tx, ty;
    bool ok;
    bilinear_taps(x, y, fw, fh, x0, y0, tx, ty, ok);
    // all taps are fetched unconditionally from clamped addresses (4 or 8 independent 128-bit loads in flight per row
    // and level); the taps outside the map are then left out of the sum like zero padding does
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    float4 t0[4], t1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int xi = min(max(x0 + (c & 1), 0), fw - 1), yi = min(max(y0 + (c >> 1), 0), fh - 1);
      const size_t texel = ((size_t)(v + 1) * fh + yi) * fw + xi;
      if (lv == 0) {
        t0[c] = __ldg(sc.img0 + texel * 2);
        t1[c] = __ldg(sc.img0 + texel * 2 + 1);
      } else {
        t0[c] = __ldg(sc.feat[lv] + texel);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int xi = x0 + (c & 1), yi = y0 + (c >> 1);
      if (!ok || xi < 0 || xi >= fw || yi < 0 || yi >= fh) continue;
      const float wgt = ((c & 1) ? tx : 1.f - tx) * ((c & 2) ? ty : 1.f - ty);
      a.x += t0[c].x * wgt; a.y += t0[c].y * wgt; a.z += t0[c].z * wgt; a.w += t0[c].w * wgt;
      if (lv == 0) { b.x += t1[c].x * wgt; b.y += t1[c].y * wgt; b.z += t1[c].z * wgt; }
    }
    if (lv == 0) {
      rec[0] = a.x; rec[1] = a.y; rec[2] = a.z; rec[3] = a.w; rec[4] = b.x; rec[5] = b.y; rec[6] = b.z;
    } else {
      rec[3 + 4 * lv] = a.x; rec[4 + 4 * lv] = a.y; rec[5 + 4 * lv] = a.z; rec[6 + 4 * lv] = a.w;
    }
    scale *= 0.5f;
  }
  rec[19] = valid ? 1.0f : 0.0f;
  rd = make_float4(ddx / dn, ddy / dn, ddz / dn, dot);
}
