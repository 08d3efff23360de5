// MatchingField.forward / depth_render (matching_field.py:18-141): the probe of the dense matching volume run for every
// pixel of a view at the stage's resolution — the upstream user of the sampler's probe.
//   k_depth_map : one warp per pixel.  Ray through the pixel (K^-1, c2w), 1 window (stage 0: [near, far]) or 2 windows
//                 (a window of width range * ratio[stage] and one of range * ratio[stage-1] around the previous stage's
//                 depth, shifted into [near, far], :101-121), n uniform depths per window (+ jitter), rank-merge of the
//                 sorted runs (== torch.sort of the concatenation), trilinear probe (grid_sample, align_corners=False,
//                 zeros), online softmax -> expected depth * cos.  Also the three per-ray sums of the occupancy
//                 regulariser (:68).
//   k_upsample_depth : F.interpolate(bilinear, align_corners=False) of the (h,w) map to the image size (:136).
// L1/L2-resident gathers (a ray's probes walk a line through the volume), instruction-issue bound like k_sample_rays.
#include "surf_internal.cuh"

#define DM_WARPS 8
#define DM_MAXS 256

// grid_sampler_3d 'bilinear', zeros padding, align_corners=False (same arithmetic as the sampler's probe, sample.cu)
__device__ __forceinline__ float dm_probe(const float* __restrict__ vol, int M, float px, float py, float pz) {
  const float ix = gs_unnorm(pz, M), iy = gs_unnorm(py, M), iz = gs_unnorm(px, M);
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const float tx = ix - fx, ty = iy - fy, tz = iz - fz;
  if (!(fx >= -1.f && fx < (float)M && fy >= -1.f && fy < (float)M && fz >= -1.f && fz < (float)M)) return 0.f;
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int xi = x0 + (c & 1), yi = y0 + ((c >> 1) & 1), zi = z0 + (c >> 2);
    if (xi < 0 || xi >= M || yi < 0 || yi >= M || zi < 0 || zi >= M) continue;
    const float w = ((c & 1) ? tx : 1.f - tx) * ((c & 2) ? ty : 1.f - ty) * ((c & 4) ? tz : 1.f - tz);
    acc += __ldg(vol + ((size_t)zi * M + yi) * M + xi) * w;
  }
  return acc;
}

__device__ __forceinline__ void dm_window(float pre_z, float range, float ratio, float near_o, float far_o, float* lo,
                                          float* hi) {
  const float sr = __fmul_rn(range, ratio);
  const float half = __fdiv_rn(sr, 2.0f);
  float n = __fsub_rn(pre_z, half), f = __fadd_rn(pre_z, half);
  if (f > far_o) n = __fsub_rn(n, __fsub_rn(f, far_o));
  if (n < near_o) f = __fadd_rn(f, __fsub_rn(near_o, n));
  *lo = fminf(fmaxf(n, near_o), far_o);
  *hi = fminf(fmaxf(f, near_o), far_o);
}

__global__ void __launch_bounds__(DM_WARPS * 32)
k_depth_map(const DevScene sc, const surf_depth_map_params p, const float* __restrict__ lin, const float* __restrict__ tx,
            const float* __restrict__ ty, const float* __restrict__ pre_depth, const float* __restrict__ t_rand,
            float* __restrict__ depth_out, float* __restrict__ occ_out) {
  __shared__ float zbuf[DM_WARPS][DM_MAXS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* zs = zbuf[wib];
  const int64_t B = (int64_t)p.h * p.w;
  const int n = p.n_samples, nw = p.n_windows, S = n * nw;
  const float range = __fsub_rn(p.far, p.near);
  for (int64_t r = (int64_t)blockIdx.x * DM_WARPS + wib; r < B; r += (int64_t)gridDim.x * DM_WARPS) {
    const float u = tx[r % p.w], v = ty[r / p.w];
    // ray of the pixel (matching_field.py:95-99)
    float cam[3], d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) cam[i] = p.Kinv[i * 3] * u + p.Kinv[i * 3 + 1] * v + p.Kinv[i * 3 + 2];
    const float nrm = sqrtf(cam[0] * cam[0] + cam[1] * cam[1] + cam[2] * cam[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) cam[i] = __fdiv_rn(cam[i], nrm);
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] = p.R[i * 3] * cam[0] + p.R[i * 3 + 1] * cam[1] + p.R[i * 3 + 2] * cam[2];
    const float camz = p.Rinv2[0] * d[0] + p.Rinv2[1] * d[1] + p.Rinv2[2] * d[2];
    // windows
    float lo[2], hi[2];
    lo[0] = p.near; hi[0] = p.far; lo[1] = p.near; hi[1] = p.far;
    if (nw == 2) {
      const float pre = pre_depth[(int64_t)(int)v * p.img_w + (int)u];
      const float pre_z = __fdiv_rn(pre, camz);
      dm_window(pre_z, range, p.ratio[0], p.near, p.far, &lo[0], &hi[0]);
      dm_window(pre_z, range, p.ratio[1], p.near, p.far, &lo[1], &hi[1]);
    }
    __syncwarp();
    for (int wi = 0; wi < nw; ++wi) {
      const float w = __fsub_rn(hi[wi], lo[wi]);
      float shift = 0.f;
      if (t_rand) shift = __fdiv_rn(__fmul_rn(__fsub_rn(t_rand[r * nw + wi], 0.5f), w), (float)n);
      for (int j = lane; j < n; j += 32) {
        float z = __fadd_rn(lo[wi], __fmul_rn(w, lin[j]));
        if (t_rand) z = __fadd_rn(z, shift);
        zs[wi * n + j] = z;
      }
    }
    __syncwarp();
    // rank of every element in the merged order (ties: window 0 first); probe; online softmax
    float m = -INFINITY, s = 0.f, t = 0.f, first6 = 0.f, dout = 0.f, nout = 0.f;
    for (int e = lane; e < S; e += 32) {
      const int a = e / n;
      const float z = zs[e];
      int rank = e - a * n;
      if (nw == 2) {
        const int bb = (1 - a) * n;
        int lo_i = 0, hi_i = n;
        while (lo_i < hi_i) {
          const int mid = (lo_i + hi_i) >> 1;
          const float q = zs[bb + mid];
          const bool before = (a == 1) ? (q <= z) : (q < z);
          if (before) lo_i = mid + 1; else hi_i = mid;
        }
        rank += lo_i;
      }
      const float px = ray_at(p.C[0], d[0], z), py = ray_at(p.C[1], d[1], z), pz = ray_at(p.C[2], d[2], z);
      const float dens = dm_probe(sc.matching, sc.mdim, px, py, pz);
      if (rank < 6) first6 += dens;
      if (sqrtf(px * px + py * py + pz * pz) > 1.0f) { dout += dens; nout += 1.f; }
      const float mn = fmaxf(m, dens);
      const float c = expf(m - mn), ex = expf(dens - mn);
      s = s * c + ex;
      t = t * c + ex * z;
      m = mn;
    }
    const float Mx = warp_max(m);
    const float c = (m == -INFINITY) ? 0.f : expf(m - Mx);
    s = warp_sum(s * c);
    t = warp_sum(t * c);
    first6 = warp_sum(first6); dout = warp_sum(dout); nout = warp_sum(nout);
    if (lane == 0) {
      depth_out[r] = __fmul_rn(__fdiv_rn(t, s), camz);
      occ_out[r * 3] = first6; occ_out[r * 3 + 1] = dout; occ_out[r * 3 + 2] = nout;
    }
    __syncwarp();
  }
}

__global__ void k_upsample_depth(const float* __restrict__ in, int h, int w, float* __restrict__ out, int H, int W) {
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  const int64_t n = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)(i / W);
    int y0, y1, x0, x1;
    float ly, lx;
    up_src(y, sh, h, &y0, &y1, &ly);
    up_src(x, sw, w, &x0, &x1, &lx);
    const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
    const float a = in[(size_t)y0 * w + x0], b = in[(size_t)y0 * w + x1], c = in[(size_t)y1 * w + x0], d = in[(size_t)y1 * w + x1];
    out[i] = __fadd_rn(__fmul_rn(hy, __fadd_rn(__fmul_rn(hx, a), __fmul_rn(lx, b))),
                       __fmul_rn(ly, __fadd_rn(__fmul_rn(hx, c), __fmul_rn(lx, d))));
  }
}

extern "C" int surf_depth_map(const surf_scene* s, const surf_depth_map_params* p, const float* d_lin, const float* d_tx,
                              const float* d_ty, const float* d_pre_depth, const float* d_t_rand, float* d_depth_lowres,
                              float* d_occ_partials, float* d_depth_full, void* stream) {
  SURF_CHECK_ARG(s && p && d_lin && d_tx && d_ty && d_depth_lowres && d_occ_partials, "null pointer");
  SURF_CHECK_ARG(s->dev.matching != nullptr, "scene has no matching volume");
  SURF_CHECK_ARG(p->n_windows == 1 || p->n_windows == 2, "n_windows must be 1 or 2");
  SURF_CHECK_ARG(p->n_samples >= 1 && p->n_samples * p->n_windows <= DM_MAXS, "too many samples per ray");
  SURF_CHECK_ARG(p->n_windows == 1 || d_pre_depth != nullptr, "two windows need the previous stage's depth map");
  SURF_CHECK_ARG(p->h >= 1 && p->w >= 1 && p->img_h >= p->h && p->img_w >= p->w, "map sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t B = (int64_t)p->h * p->w;
  const int64_t blocks = (B + DM_WARPS - 1) / DM_WARPS;
  const int64_t cap = (int64_t)surf_num_sms() * 8;
  k_depth_map<<<(int)(blocks < cap ? blocks : cap), DM_WARPS * 32, 0, st>>>(s->dev, *p, d_lin, d_tx, d_ty, d_pre_depth,
                                                                           d_t_rand, d_depth_lowres, d_occ_partials);
  SURF_LAUNCH_CHECK();
  if (d_depth_full) {
    const int64_t n = (int64_t)p->img_h * p->img_w;
    const int64_t b2 = (n + 255) / 256;
    k_upsample_depth<<<(int)(b2 < cap ? b2 : cap), 256, 0, st>>>(d_depth_lowres, p->h, p->w, d_depth_full, p->img_h,
                                                                p->img_w);
    SURF_LAUNCH_CHECK();
  }
  return 0;
}
