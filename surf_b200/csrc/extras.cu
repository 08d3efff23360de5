// Training-only tail of render_core (implicit_surface.py:218-245) and surface_patch_warp2 / patch_homography
// (projector.py:560-645): the zero-crossing surface point of every ray, its unit normal in the reference camera frame
// (a third SDF-MLP pass on B points), and the plane-induced homography warp of an 11x11 pixel patch of the 12-channel
// feature maps into every source view.  HBM/L2-bound gathers; tiny next to the render itself (B rays x 121 pixels
// x (1 + V) views x 4 taps x 48 B).
//
//   k_build_warp12   : the 12-channel maps the loss compares (implicit_surface.py:229-241) = level-0 features |
//                      level-1 | level-2 features up-sampled to the level-0 size (F.interpolate bilinear,
//                      align_corners=False), NHWC [nv][H][W][12] so that a tap is three 128-bit loads.  Built once per
//                      surf_scene_set_views (the reference re-interpolates the full maps on every render call).
//   k_sdf0_points    : z_sdf0 clamped to [0, max z_vals of the call] (Q15) -> pts_sdf0 = o + d * z.
//   k_patch_warp     : one thread per (view, ray, patch pixel): normal = R0^T (g / |g|), homography
//                      H = K_src (R_rel + (R_src (C_ref - C_src)) n^T / (n . p_ref + 1e-10)) K_ref^-1, warped pixel
//                      = H (u, v, 1) / (w + 1e-8), grid_sample(bilinear, zeros, align_corners=True).
#include "surf_internal.cuh"

__device__ __forceinline__ float4 up_tap(const float4* __restrict__ f, int w, int y0, int y1, int x0, int x1, float ly,
                                         float lx) {
  const float4 a = __ldg(f + (size_t)y0 * w + x0), b = __ldg(f + (size_t)y0 * w + x1);
  const float4 c = __ldg(f + (size_t)y1 * w + x0), d = __ldg(f + (size_t)y1 * w + x1);
  const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
  float4 r;
  // ATen: w_h0 * (w_w0 * v00 + w_w1 * v01) + w_h1 * (w_w0 * v10 + w_w1 * v11)
  r.x = __fadd_rn(__fmul_rn(hy, __fadd_rn(__fmul_rn(hx, a.x), __fmul_rn(lx, b.x))), __fmul_rn(ly, __fadd_rn(__fmul_rn(hx, c.x), __fmul_rn(lx, d.x))));
  r.y = __fadd_rn(__fmul_rn(hy, __fadd_rn(__fmul_rn(hx, a.y), __fmul_rn(lx, b.y))), __fmul_rn(ly, __fadd_rn(__fmul_rn(hx, c.y), __fmul_rn(lx, d.y))));
  r.z = __fadd_rn(__fmul_rn(hy, __fadd_rn(__fmul_rn(hx, a.z), __fmul_rn(lx, b.z))), __fmul_rn(ly, __fadd_rn(__fmul_rn(hx, c.z), __fmul_rn(lx, d.z))));
  r.w = __fadd_rn(__fmul_rn(hy, __fadd_rn(__fmul_rn(hx, a.w), __fmul_rn(lx, b.w))), __fmul_rn(ly, __fadd_rn(__fmul_rn(hx, c.w), __fmul_rn(lx, d.w))));
  return r;
}

__global__ void k_build_warp12(const DevScene sc, float4* __restrict__ out) {
  const int64_t n = (int64_t)sc.nv * sc.H * sc.W;
  const float sh1 = (float)sc.fh[1] / (float)sc.H, sw1 = (float)sc.fw[1] / (float)sc.W;
  const float sh2 = (float)sc.fh[2] / (float)sc.H, sw2 = (float)sc.fw[2] / (float)sc.W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % sc.W);
    const int y = (int)((i / sc.W) % sc.H);
    const int v = (int)(i / ((int64_t)sc.W * sc.H));
    const float4 t0 = __ldg(sc.img0 + i * 2), t1 = __ldg(sc.img0 + i * 2 + 1);
    out[i * 3] = make_float4(t0.w, t1.x, t1.y, t1.z);
    int y0, y1, x0, x1;
    float ly, lx;
    up_src(y, sh1, sc.fh[1], &y0, &y1, &ly);
    up_src(x, sw1, sc.fw[1], &x0, &x1, &lx);
    out[i * 3 + 1] = up_tap(sc.feat[1] + (size_t)v * sc.fh[1] * sc.fw[1], sc.fw[1], y0, y1, x0, x1, ly, lx);
    up_src(y, sh2, sc.fh[2], &y0, &y1, &ly);
    up_src(x, sw2, sc.fw[2], &x0, &x1, &lx);
    out[i * 3 + 2] = up_tap(sc.feat[2] + (size_t)v * sc.fh[2] * sc.fw[2], sc.fw[2], y0, y1, x0, x1, ly, lx);
  }
}

__global__ void k_sdf0_points(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                              const float* __restrict__ z_cross, const float* __restrict__ z_max, int64_t B,
                              float* __restrict__ pts) {
  const float zm = *z_max;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    float z = z_cross[r];
    if (z < 0.f) z = 0.f;          // torch.where(z < 0, 0, z): NaN stays NaN like the reference
    if (z > zm) z = 0.f;
    pts[r * 3] = ray_at(rays_o[r * 3], rays_d[r * 3], z);
    pts[r * 3 + 1] = ray_at(rays_o[r * 3 + 1], rays_d[r * 3 + 1], z);
    pts[r * 3 + 2] = ray_at(rays_o[r * 3 + 2], rays_d[r * 3 + 2], z);
  }
}

// grid_sample(bilinear, padding zeros, align_corners=True) of the 12-channel NHWC map of one view at pixel (x, y)
__device__ __forceinline__ void sample12(const float4* __restrict__ img, int H, int W, float x, float y, float4 (&o)[3]) {
  // the reference normalises the pixel to [-1,1] and grid_sample un-normalises it again (projector.py:609-612)
  const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, x), (float)(W - 1)), 1.0f);
  const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, y), (float)(H - 1)), 1.0f);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = __fsub_rn(ix, x0f), wx0 = __fsub_rn(x0f + 1.0f, ix);
  const float wy1 = __fsub_rn(iy, y0f), wy0 = __fsub_rn(y0f + 1.0f, iy);
  o[0] = o[1] = o[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  // NaN / inf coordinates: every comparison below is false -> zeros, like ATen's within_bounds_2d on the converted index
  if (!(ix > -2.0f && ix < (float)W + 1.0f && iy > -2.0f && iy < (float)H + 1.0f)) return;
  const int x0 = (int)x0f, y0 = (int)y0f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {      // nw, ne, sw, se (ATen order)
    const int xi = x0 + (t & 1), yi = y0 + (t >> 1);
    if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;
    const float w = __fmul_rn((t & 1) ? wx1 : wx0, (t >> 1) ? wy1 : wy0);
    const float4* p = img + ((size_t)yi * W + xi) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 v = __ldg(p + c);
      o[c].x += v.x * w; o[c].y += v.y * w; o[c].z += v.z * w; o[c].w += v.w * w;
    }
  }
}

__device__ __forceinline__ void mat3_mul(const float* A, const float* Bm, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * Bm[j] + A[i * 3 + 1] * Bm[3 + j] + A[i * 3 + 2] * Bm[6 + j];
}

__global__ void __launch_bounds__(128)
k_patch_warp(const surf_extras_params p, const float4* __restrict__ warp12, int H, int W, const float* __restrict__ pts,
             const float* __restrict__ grad, int64_t B, float* __restrict__ normal_out, float4* __restrict__ ref_val,
             float4* __restrict__ src_val) {
  const int ps = p.patch_size, hp = ps / 2, npx = ps * ps;
  const int nvw = p.n_src + 1;
  const int64_t total = (int64_t)nvw * B * npx;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % npx);
    const int64_t b = (i / npx) % B;
    const int vv = (int)(i / ((int64_t)npx * B));       // 0 = reference view, 1.. = source views
    // unit normal in the reference camera frame (implicit_surface.py:223-227)
    float g[3] = {grad[b * 3], grad[b * 3 + 1], grad[b * 3 + 2]};
    float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    if (nrm <= 0.f) nrm = 1e-8f;
    g[0] = __fdiv_rn(g[0], nrm); g[1] = __fdiv_rn(g[1], nrm); g[2] = __fdiv_rn(g[2], nrm);
    float n[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) n[r] = p.R0t[r * 3] * g[0] + p.R0t[r * 3 + 1] * g[1] + p.R0t[r * 3 + 2] * g[2];
    if (vv == 0 && k == 0 && normal_out) { normal_out[b * 3] = n[0]; normal_out[b * 3 + 1] = n[1]; normal_out[b * 3 + 2] = n[2]; }
    // the point in the reference camera frame and its pixel (projector.py:573-579, 598-600)
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      pc[r] = __fadd_rn(p.R0t[r * 3] * pts[b * 3] + p.R0t[r * 3 + 1] * pts[b * 3 + 1] + p.R0t[r * 3 + 2] * pts[b * 3 + 2], p.t0[r]);
    float pr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) pr[r] = p.K0[r * 3] * pc[0] + p.K0[r * 3 + 1] * pc[1] + p.K0[r * 3 + 2] * pc[2];
    const float px = __fdiv_rn(pr[0], __fadd_rn(pr[2], 1e-8f)), py = __fdiv_rn(pr[1], __fadd_rn(pr[2], 1e-8f));
    const float u = __fadd_rn(px, (float)(k % ps - hp)), v = __fadd_rn(py, (float)(k / ps - hp));
    float4 o[3];
    if (vv == 0) {
      sample12(warp12, H, W, u, v, o);
      float4* dst = ref_val + ((size_t)b * npx + k) * 3;
      dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    } else {
      const int s = vv - 1;
      const float disp = __fadd_rn(n[0] * pc[0] + n[1] * pc[1] + n[2] * pc[2], 1e-10f);
      float M[9], A[9], Hm[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) M[r * 3 + c] = __fadd_rn(p.Rrel[s][r * 3 + c], __fdiv_rn(__fmul_rn(p.RC[s][r], n[c]), disp));
      mat3_mul(p.Ksrc[s], M, A);
      mat3_mul(A, p.K0inv, Hm);
      const float tx = Hm[0] * u + Hm[1] * v + Hm[2], ty = Hm[3] * u + Hm[4] * v + Hm[5], tz = Hm[6] * u + Hm[7] * v + Hm[8];
      const float wx = __fdiv_rn(tx, __fadd_rn(tz, 1e-8f)), wy = __fdiv_rn(ty, __fadd_rn(tz, 1e-8f));
      sample12(warp12 + (size_t)vv * H * W * 3, H, W, wx, wy, o);
      float4* dst = src_val + (((size_t)s * B + b) * npx + k) * 3;
      dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    }
  }
}

extern "C" size_t surf_extras_workspace_bytes(int64_t n_rays) { return (size_t)n_rays * 4 * sizeof(float) + 256; }

extern "C" int surf_render_extras(surf_scene* s, const surf_net* n, const surf_extras_params* p, const float* d_rays_o,
                                  const float* d_rays_d, const float* d_z_cross, const float* d_z_max, int64_t n_rays,
                                  float* d_pts_sdf0, float* d_normal_sdf0, float* d_ref_val, float* d_src_val,
                                  void* d_workspace, size_t workspace_bytes, int32_t mlp_mode, void* stream) {
  SURF_CHECK_ARG(s && n && p && d_rays_o && d_rays_d && d_z_cross && d_z_max && d_pts_sdf0 && d_ref_val && d_src_val &&
                     d_workspace, "null pointer");
  if (n_rays <= 0) return 0;
  SURF_CHECK_ARG(workspace_bytes >= surf_extras_workspace_bytes(n_rays), "workspace too small");
  SURF_CHECK_ARG(p->n_src == s->dev.V, "n_src differs from the scene's source views");
  SURF_CHECK_ARG(p->patch_size >= 1 && (p->patch_size & 1), "patch_size must be odd");
  SURF_CHECK_ARG(s->dev.img0 && s->dev.feat[1] && s->dev.feat[2], "scene has no views (surf_scene_set_views)");
  cudaStream_t st = (cudaStream_t)stream;
  const DevScene& d = s->dev;
  // the 12-channel warp maps of the current views (rebuilt when surf_scene_set_views ran since)
  const size_t wbytes = (size_t)d.nv * d.H * d.W * 48;
  if (s->warp_bytes != wbytes) {
    if (s->warp12) SURF_CUDA(cudaFreeAsync(s->warp12, st));
    s->warp12 = nullptr;
    s->warp_bytes = 0;
    SURF_CUDA(cudaMallocAsync(&s->warp12, wbytes, st));
    s->warp_bytes = wbytes;
    s->warp_version = -1;
  }
  if (s->warp_version != s->views_version) {
    const int64_t npix = (int64_t)d.nv * d.H * d.W;
    const int64_t cap = (int64_t)surf_num_sms() * 16;
    const int64_t blocks = (npix + 255) / 256;
    k_build_warp12<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(d, (float4*)s->warp12);
    SURF_LAUNCH_CHECK();
    s->warp_version = s->views_version;
  }
  {
    const int64_t blocks = (n_rays + 255) / 256;
    k_sdf0_points<<<(int)(blocks < 1024 ? blocks : 1024), 256, 0, st>>>(d_rays_o, d_rays_d, d_z_cross, d_z_max, n_rays,
                                                                         d_pts_sdf0);
    SURF_LAUNCH_CHECK();
  }
  // third SDF-MLP pass: gradient at the surface points (implicit_surface.py:222)
  float* w_sdf = (float*)d_workspace;
  float* w_grad = w_sdf + ((n_rays + 63) / 64) * 64;
  PointSource src = {};
  src.mode = 0;
  src.pts = d_pts_sdf0;
  src.n = n_rays;
  int rc = launch_sdf_mlp(s, n, src, w_sdf, w_grad, false, mlp_mode, st);
  if (rc) return rc;
  {
    const int64_t total = (int64_t)(p->n_src + 1) * n_rays * p->patch_size * p->patch_size;
    const int64_t cap = (int64_t)surf_num_sms() * 16;
    const int64_t blocks = (total + 127) / 128;
    k_patch_warp<<<(int)(blocks < cap ? blocks : cap), 128, 0, st>>>(*p, (const float4*)s->warp12, d.H, d.W, d_pts_sdf0,
                                                                    w_grad, n_rays, d_normal_sdf0, (float4*)d_ref_val,
                                                                    (float4*)d_src_val);
    SURF_LAUNCH_CHECK();
  }
  return 0;
}
