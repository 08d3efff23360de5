// Second-order term of SDFNetworkSparse.gradient (sdf_network.py:129-152): next to d sdf / d x the reference returns
//   smooth = d/dx [ sum_j d sdf / d x_j ] = Hessian(sdf) . (1,1,1)
// (double autograd), which render_core reduces to `smooth_error` (implicit_surface.py:172).  Training-only: validate()
// never reads it, so it lives in its own plain fp32 kernel instead of the tensor-core render path.
//
// Analytic form: forward-mode tangent along v = (1,1,1) through the reverse pass ("forward over reverse").  With
// a_l the input of lin_l, z_l = W_l a_l + b_l, h_l = softplus_100(z_l):
//   forward      : z_l, and the tangent zd_l = W_l ad_l, hd_l = s'(z_l) zd_l      (ad_0 = d PE / d eps, fd = J_feat v)
//   reverse      : delta_l = ga_{l+1}[hidden] * s'(z_l),  ga_l = W_l^T delta_l
//   reverse, dot : deltad_l = gad_{l+1}[hidden] * s'(z_l) + ga_{l+1}[hidden] * s''(z_l) zd_l,  gad_l = W_l^T deltad_l
//   smooth       = scale (Jd_pe^T g_pe + J_pe^T gd_pe) + Jd_feat^T g_feat + J_feat^T gd_feat
// where g_pe / g_feat collect the PE (lin0 + skip layer) and feature columns of ga, gd_* those of gad, Jd_pe is the
// derivative of the PE Jacobian along v (-f^2 sin / -f^2 cos) and Jd_feat the mixed second derivatives of the
// trilinear interpolant (the pure second derivatives vanish).
// 16 points per 256-thread block (thread = (neuron, point half): 8 points each, so a weight read from L2 feeds 16
// points), thread = output neuron (forward) / input column (reverse), everything of the 16 points
// in shared memory, weights read from L2 (k-major for the forward, row-major for the reverse: both coalesced).
#include <vector>

#include "smooth_common.cuh"
#include "surf_internal.cuh"

#define SM_NP 16             // points per block iteration
#define SM_PH 8              // points per thread (two thread halves of 128 share the weights)
#define SM_THREADS 256
#define SM_PS 20              // row stride of the k-major operands
#define SM_STRIDE 160

struct SmoothSmem {
  // GEMV operands k-major with the 16 points innermost (row stride SM_PS floats, 16-byte aligned): a thread reads its 8
  // points of one k as two 128-bit broadcast loads (a [point][k] layout costs 16 scalar broadcasts per k and made the
  // kernel shared-memory-bandwidth bound: 5.6 ms for 70 k points)
  // (the reverse pass needs softplus'(z_l) and softplus''(z_l) zd_l of every layer: 96 KB per 16 points.  They go to
  // a per-block L2-resident scratch, written and read by the same thread, so that three blocks fit on an SM)
  float A[SM_STRIDE][SM_PS], Ad[SM_STRIDE][SM_PS];       // forward operands; the reverse pass reuses them as ga / gad
#define GA A
#define GAd Ad
  float D[128][SM_PS], Dd[128][SM_PS];
  float PE[SM_NP][28], PEd[SM_NP][28], FT[SM_NP][28], FTd[SM_NP][28];
  float gpe[SM_NP][28], gdpe[SM_NP][28], gft[SM_NP][28], gdft[SM_NP][28];
  float pt[SM_NP][3];
  float part[SM_NP][4][3];
};

__device__ __forceinline__ void sm_softplus(float z, float& h, float& d1, float& d2) {
  // softplus(beta = 100): h, h' = sigmoid(100 z), h'' = 100 h' (1 - h')
  const float t = 100.0f * z;
  const float e = expf(-fabsf(t));
  h = fmaxf(z, 0.f) + log1pf(e) * 0.01f;
  const float r = 1.0f / (1.0f + e);
  d1 = t >= 0.f ? r : 1.0f - r;
  d2 = 100.0f * e * r * r;
}

#define SM_BLOCKS_PER_SM 3
#define SM_SCRATCH_FLOAT2 (6 * SM_NP * 128)     // per block: (softplus', softplus'' zd) per (layer, point, neuron)

__global__ void __launch_bounds__(SM_THREADS, SM_BLOCKS_PER_SM)
k_sdf_smooth(const DevScene sc, const DevNet net, const float* __restrict__ wk, const int* __restrict__ wk_off,
             const float* __restrict__ wr, const int* __restrict__ wr_off, const float* __restrict__ pts,
             const uint8_t* __restrict__ flags, int64_t n, float* __restrict__ grad_out, float* __restrict__ smooth_out,
             float2* __restrict__ scratch_all) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SmoothSmem& S = *reinterpret_cast<SmoothSmem*>(smem_raw);
  const int tid = threadIdx.x;
  float2* scratch = scratch_all + (size_t)blockIdx.x * SM_SCRATCH_FLOAT2;
  const int nid = tid & 127, pb = (tid >> 7) * SM_PH;     // my neuron / column, my first point
  const int pe_dim = net.pe_dim;       // 27
  for (int64_t base = (int64_t)blockIdx.x * SM_NP; base < n; base += (int64_t)gridDim.x * SM_NP) {
    // a group of points none of which is evaluated (masked-out samples of a ray): zeros, no work
    int live = (flags == nullptr) ? 1 : 0;
    if (flags != nullptr && tid < SM_NP && base + tid < n) live = (flags[base + tid] >> 1) & 1;
    if (!__syncthreads_or(live)) {
      if (tid < SM_NP * 3 && base + tid / 3 < n) {
        smooth_out[(base + tid / 3) * 3 + tid % 3] = 0.f;
        if (grad_out) grad_out[(base + tid / 3) * 3 + tid % 3] = 0.f;
      }
      continue;
    }
    // ---- inputs: features + tangent (thread = (point, level)), positional encoding + tangent (thread = point) ----
    if (tid < SM_NP * 4) {
      const int p = tid >> 2, lv = tid & 3;
      const int64_t i = base + p;
      float f7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, fd7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (i < n && lv < sc.n_levels) sparse_value_tangent(sc, lv, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], f7, fd7);
#pragma unroll
      for (int c = 0; c < 7; ++c) { S.FT[p][lv * 7 + c] = f7[c]; S.FTd[p][lv * 7 + c] = fd7[c]; }
    } else if (tid < SM_NP * 5) {
      const int p = tid - SM_NP * 4;
      const int64_t i = base + p;
      float x[3] = {0.f, 0.f, 0.f};
      if (i < n) { x[0] = pts[i * 3]; x[1] = pts[i * 3 + 1]; x[2] = pts[i * 3 + 2]; }
      for (int d = 0; d < 3; ++d) {
        S.pt[p][d] = x[d];
        const float X = x[d] * net.scale;
        S.PE[p][d] = X;
        S.PEd[p][d] = net.scale;                      // d X / d eps, v = (1,1,1)
        float fr = 1.0f;
        for (int f = 0; f < net.multires; ++f) {
          float sn, cs;
          sincosf(X * fr, &sn, &cs);
          S.PE[p][3 + 6 * f + d] = sn;      S.PEd[p][3 + 6 * f + d] = fr * net.scale * cs;
          S.PE[p][3 + 6 * f + 3 + d] = cs;  S.PEd[p][3 + 6 * f + 3 + d] = -fr * net.scale * sn;
          fr *= 2.0f;
        }
      }
    }
    for (int i = tid; i < SM_NP * 28; i += SM_THREADS) {
      (&S.gpe[0][0])[i] = 0.f; (&S.gdpe[0][0])[i] = 0.f; (&S.gft[0][0])[i] = 0.f; (&S.gdft[0][0])[i] = 0.f;
    }
    __syncthreads();
    for (int i = tid; i < SM_NP * pe_dim; i += SM_THREADS) {
      S.A[i % pe_dim][i / pe_dim] = S.PE[i / pe_dim][i % pe_dim];
      S.Ad[i % pe_dim][i / pe_dim] = S.PEd[i / pe_dim][i % pe_dim];
    }
    __syncthreads();
    // ---- forward with tangent ----
    int in_dim = pe_dim;
    for (int l = 0; l < 6; ++l) {
      const int od = net.out_dim[l];
      const float* W = wk + wk_off[l];
      float acc[SM_PH], accd[SM_PH];
#pragma unroll
      for (int p = 0; p < SM_PH; ++p) { acc[p] = 0.f; accd[p] = 0.f; }
      if (nid < od) {
        // eight weight loads in flight per thread (the weights come from L2: one block per SM, 8 warps — a load per
        // iteration leaves the warp waiting on L2 latency)
        for (int k0 = 0; k0 < in_dim; k0 += 8) {
          float w8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w8[j] = (k0 + j < in_dim) ? __ldg(W + (size_t)(k0 + j) * SM_STRIDE + nid) : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = (k0 + j < in_dim) ? k0 + j : 0;
            const float w = w8[j];
            const float4 a0 = *reinterpret_cast<const float4*>(&S.A[k][pb]), a1 = *reinterpret_cast<const float4*>(&S.A[k][pb + 4]);
            const float4 d0 = *reinterpret_cast<const float4*>(&S.Ad[k][pb]), d1 = *reinterpret_cast<const float4*>(&S.Ad[k][pb + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int p = 0; p < SM_PH; ++p) { acc[p] = fmaf(av[p], w, acc[p]); accd[p] = fmaf(dv[p], w, accd[p]); }
          }
        }
        const float b = W[(size_t)in_dim * SM_STRIDE + nid];
#pragma unroll
        for (int p = 0; p < SM_PH; ++p) acc[p] += b;
      }
      __syncthreads();          // every thread is done reading A / Ad of this layer
      // next input [h | PE at the skip layer | features] and its tangent; the derivatives the reverse pass needs
#pragma unroll
      for (int p = 0; p < SM_PH; ++p) {
        float h = 0.f, hd = 0.f;
        if (nid < od) {
          float d1, d2;
          sm_softplus(acc[p], h, d1, d2);
          hd = d1 * accd[p];
          scratch[(size_t)(l * SM_NP + pb + p) * 128 + nid] = make_float2(d1, d2 * accd[p]);
        } else if (l + 1 == net.skip_layer && nid < od + pe_dim) {
          h = S.PE[pb + p][nid - od]; hd = S.PEd[pb + p][nid - od];
        }
        S.A[nid][pb + p] = h; S.Ad[nid][pb + p] = hd;
      }
      for (int i = tid; i < SM_NP * 28; i += SM_THREADS) {
        S.A[128 + i % 28][i / 28] = S.FT[i / 28][i % 28];
        S.Ad[128 + i % 28][i / 28] = S.FTd[i / 28][i % 28];
      }
      in_dim = 128 + 28;
      __syncthreads();
    }
    // ---- reverse with tangent: ga_6 = row 0 of lin6 / scale, gad_6 = 0 ----
    // Work split of one reverse layer (156 input columns, od rows): thread nid owns hidden column nid over all rows; the
    // 28 feature columns (128 + jq) are split by ROW QUARTER oq = nid >> 5 and their partial sums stay in registers
    // across the layers (g_feat is the sum over the layers anyway) — a "column 128 + nid for nid < 28" split keeps one
    // warp busy twice as long as the other three and was 17 % of the kernel's stall samples at the layer barrier.
    // lin0 (27 PE columns) is split over the row quarters the same way.
    const int jq = nid & 31, oq = nid >> 5;
    float gf[SM_PH], gfd[SM_PH];
#pragma unroll
    for (int p = 0; p < SM_PH; ++p) { gf[p] = 0.f; gfd[p] = 0.f; }
    {
      const float* W6 = wr + wr_off[6];
      const float w = W6[nid] * net.inv_scale;
#pragma unroll
      for (int p = 0; p < SM_PH; ++p) { S.GA[nid][pb + p] = w; S.GAd[nid][pb + p] = 0.f; }
    }
    __syncthreads();
    for (int l = 5; l >= 0; --l) {
      const int od = net.out_dim[l];
      // PE columns of the skip layer's input
      if (l + 1 == net.skip_layer)
        for (int i = tid; i < SM_NP * pe_dim; i += SM_THREADS) {
          S.gpe[i / pe_dim][i % pe_dim] += S.GA[od + i % pe_dim][i / pe_dim];
          S.gdpe[i / pe_dim][i % pe_dim] += S.GAd[od + i % pe_dim][i / pe_dim];
        }
      // delta_l and its tangent
#pragma unroll
      for (int p = 0; p < SM_PH; ++p) {
        float dl = 0.f, dd = 0.f;
        if (nid < od) {
          const float2 d = scratch[(size_t)(l * SM_NP + pb + p) * 128 + nid];     // (s'(z_l), s''(z_l) zd_l)
          const float ga = S.GA[nid][pb + p];
          dl = ga * d.x;
          dd = fmaf(S.GAd[nid][pb + p], d.x, ga * d.y);
        }
        S.D[nid][pb + p] = dl; S.Dd[nid][pb + p] = dd;
      }
      __syncthreads();
      const float* W = wr + wr_off[l];
      // sum over rows [o_lo, o_hi) of W[o][col] * (D[o], Dd[o]) for my 8 points, eight weight loads in flight
      auto column = [&](int col, int o_lo, int o_hi, float (&ga)[SM_PH], float (&gad)[SM_PH]) {
        for (int o0 = o_lo; o0 < o_hi; o0 += 8) {
          float w8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w8[j] = (o0 + j < o_hi) ? __ldg(W + (size_t)(o0 + j) * SM_STRIDE + col) : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int o = (o0 + j < o_hi) ? o0 + j : o_lo;
            const float w = w8[j];
            const float4 a0 = *reinterpret_cast<const float4*>(&S.D[o][pb]), a1 = *reinterpret_cast<const float4*>(&S.D[o][pb + 4]);
            const float4 d0 = *reinterpret_cast<const float4*>(&S.Dd[o][pb]), d1 = *reinterpret_cast<const float4*>(&S.Dd[o][pb + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int p = 0; p < SM_PH; ++p) { ga[p] = fmaf(av[p], w, ga[p]); gad[p] = fmaf(dv[p], w, gad[p]); }
          }
        }
      };
      const int q_lo = oq * 32 < od ? oq * 32 : od, q_hi = oq * 32 + 32 < od ? oq * 32 + 32 : od;
      if (l > 0) {
        float ga[SM_PH], gad[SM_PH];
#pragma unroll
        for (int p = 0; p < SM_PH; ++p) { ga[p] = 0.f; gad[p] = 0.f; }
        column(nid, 0, od, ga, gad);
#pragma unroll
        for (int p = 0; p < SM_PH; ++p) { S.GA[nid][pb + p] = ga[p]; S.GAd[nid][pb + p] = gad[p]; }
        if (jq < 28) column(128 + jq, q_lo, q_hi, gf, gfd);
      } else {
        // lin0: partial sums of PE column jq over my row quarter -> GA rows oq * 32 + jq, reduced below
        float ga[SM_PH], gad[SM_PH];
#pragma unroll
        for (int p = 0; p < SM_PH; ++p) { ga[p] = 0.f; gad[p] = 0.f; }
        if (jq < pe_dim) column(jq, q_lo, q_hi, ga, gad);
#pragma unroll
        for (int p = 0; p < SM_PH; ++p) { S.GA[nid][pb + p] = ga[p]; S.GAd[nid][pb + p] = gad[p]; }
      }
      __syncthreads();
    }
    // the feature-column partial sums of the four row quarters (D / Dd are free now)
#pragma unroll
    for (int p = 0; p < SM_PH; ++p) { S.D[nid][pb + p] = gf[p]; S.Dd[nid][pb + p] = gfd[p]; }
    __syncthreads();
    {
      const float* W6 = wr + wr_off[6];
      for (int i = tid; i < SM_NP * 28; i += SM_THREADS) {
        const int p = i / 28, j = i % 28;
        S.gft[p][j] = W6[128 + j] * net.inv_scale + ((S.D[j][p] + S.D[32 + j][p]) + (S.D[64 + j][p] + S.D[96 + j][p]));
        S.gdft[p][j] = (S.Dd[j][p] + S.Dd[32 + j][p]) + (S.Dd[64 + j][p] + S.Dd[96 + j][p]);
      }
      // lin0's input is the positional encoding
      for (int i = tid; i < SM_NP * pe_dim; i += SM_THREADS) {
        const int p = i / pe_dim, j = i % pe_dim;
        S.gpe[p][j] += (S.GA[j][p] + S.GA[32 + j][p]) + (S.GA[64 + j][p] + S.GA[96 + j][p]);
        S.gdpe[p][j] += (S.GAd[j][p] + S.GAd[32 + j][p]) + (S.GAd[64 + j][p] + S.GAd[96 + j][p]);
      }
    }
    __syncthreads();
    // ---- d/dx: feature part per (point, level), PE part per point ----
    if (tid < SM_NP * 4) {
      const int p = tid >> 2, lv = tid & 3;
      const int64_t i = base + p;
      float o3[3] = {0.f, 0.f, 0.f}, od3[3] = {0.f, 0.f, 0.f}, om3[3] = {0.f, 0.f, 0.f};
      if (i < n && lv < sc.n_levels) {
        const float zero7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float t3[3];
        // J^T gd (scaled 1/vs): a first-order pass with gd;  J^T g and the mixed terms (1/vs^2): a pass with (g, 0)
        sparse_back_tangent(sc, lv, S.pt[p][0], S.pt[p][1], S.pt[p][2], &S.gdft[p][lv * 7], zero7, od3, t3);
        sparse_back_tangent(sc, lv, S.pt[p][0], S.pt[p][1], S.pt[p][2], &S.gft[p][lv * 7], zero7, o3, om3);
        const float inv = 1.0f / sc.voxel[lv];
        om3[0] *= inv * inv; om3[1] *= inv * inv; om3[2] *= inv * inv;
      }
      for (int d = 0; d < 3; ++d) S.part[p][lv][d] = od3[d] + om3[d];
      if (grad_out) for (int d = 0; d < 3; ++d) S.D[lv * 3 + d][p] = o3[d];
    }
    __syncthreads();
    if (tid < SM_NP * 3) {
      const int p = tid / 3, d = tid % 3;
      const int64_t i = base + p;
      if (i < n) {
        const float X = S.pt[p][d] * net.scale;
        float g1 = S.gpe[p][d];                 // J_pe^T g_pe (first order, per unit X)
        float s2 = S.gdpe[p][d];                // J_pe^T gd_pe + Jd_pe^T g_pe
        float fr = 1.0f;
        for (int f = 0; f < net.multires; ++f) {
          const float sn = S.PE[p][3 + 6 * f + d], cs = S.PE[p][3 + 6 * f + 3 + d];
          const float gs = S.gpe[p][3 + 6 * f + d], gc = S.gpe[p][3 + 6 * f + 3 + d];
          g1 += fr * (gs * cs - gc * sn);
          s2 += fr * (S.gdpe[p][3 + 6 * f + d] * cs - S.gdpe[p][3 + 6 * f + 3 + d] * sn);
          s2 -= fr * fr * net.scale * (gs * sn + gc * cs);     // d/d eps of (f cos, -f sin) along dX = scale
          fr *= 2.0f;
        }
        (void)X;
        const float feat2 = S.part[p][0][d] + S.part[p][1][d] + S.part[p][2][d] + S.part[p][3][d];
        const bool on = flags == nullptr || ((flags[i] >> 1) & 1);
        smooth_out[i * 3 + d] = on ? fmaf(s2, net.scale, feat2) : 0.f;
        if (grad_out) {
          const float feat1 = S.D[d][p] + S.D[3 + d][p] + S.D[6 + d][p] + S.D[9 + d][p];
          grad_out[i * 3 + d] = on ? fmaf(g1, net.scale, feat1) : 0.f;
        }
      }
    }
  }
}

// row-major fp32 copies of lin0..lin6 (stride 160) for the reverse passes of k_sdf_smooth
int surf_build_smooth_weights(const std::vector<std::vector<float>>& W, const surf_net_inputs* in, surf_net* net,
                              cudaStream_t st, int (*dev_alloc)(surf_net*, void**, size_t)) {
  std::vector<float> wr;
  std::vector<int> off(8, 0);
  for (int l = 0; l < 7; ++l) {
    const int O = in->out_dim[l], I = in->in_dim[l];
    if (I > SM_STRIDE || (l < 6 && O > 128)) {
      surf_set_error("lin%d: %d x %d unsupported by the second-order kernel", l, O, I);
      return -1;
    }
    off[l] = (int)wr.size();
    const int rows = (l == 6) ? 1 : O;
    wr.resize(wr.size() + (size_t)rows * SM_STRIDE, 0.f);
    for (int o = 0; o < rows; ++o)
      for (int k = 0; k < I; ++k) wr[off[l] + (size_t)o * SM_STRIDE + k] = W[l][(size_t)o * I + k];
  }
  void* p = nullptr;
  int rc = dev_alloc(net, &p, wr.size() * sizeof(float));
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, wr.data(), wr.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  net->w_rows = (const float*)p;
  rc = dev_alloc(net, &p, off.size() * sizeof(int));
  if (rc) return rc;
  SURF_CUDA(cudaMemcpyAsync(p, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  net->w_rows_off = (const int*)p;
  SURF_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int surf_sdf_smooth(const surf_scene* s, const surf_net* n, const float* d_pts, int64_t n_pts,
                               const uint8_t* d_flags, float* d_grad, float* d_smooth, int32_t mlp_mode, void* stream) {
  if (n_pts <= 0) return 0;
  SURF_CHECK_ARG(s && n && d_pts && d_smooth, "null pointer");
  SURF_CHECK_ARG(mlp_mode == SURF_MLP_FFMA || mlp_mode == SURF_MLP_TC || mlp_mode == SURF_MLP_TC_FAST, "mlp_mode must be SURF_MLP_FFMA, SURF_MLP_TC or SURF_MLP_TC_FAST");
  if (mlp_mode != SURF_MLP_FFMA && n->tc_ok)
    return launch_sdf_smooth_tc(s, n, d_pts, n_pts, d_flags, d_grad, d_smooth, mlp_mode == SURF_MLP_TC_FAST, (cudaStream_t)stream);
  SURF_CHECK_ARG(n->w_rows && n->w_full, "network without second-order weights");
  SURF_CHECK_ARG(n->dev.pe_dim <= 28 && n->dev.out_dim[6] >= 1, "unsupported network shape");
  int rc = surf_ensure_dyn_smem((const void*)k_sdf_smooth, (int)sizeof(SmoothSmem));
  if (rc) return rc;
  const int64_t blocks = (n_pts + SM_NP - 1) / SM_NP;
  int64_t cap = (int64_t)n->n_sm * SM_BLOCKS_PER_SM;
  // the per-block derivative scratch lives in the network's scratch buffer (shared with the FFMA reverse pass: calls
  // on one network handle are stream-ordered)
  const int64_t fit = (int64_t)(n->scratch_bytes / (SM_SCRATCH_FLOAT2 * sizeof(float2)));
  if (cap > fit) cap = fit;
  SURF_CHECK_ARG(cap >= 1 && n->scratch, "network without scratch buffer");
  k_sdf_smooth<<<(int)(blocks < cap ? blocks : cap), SM_THREADS, sizeof(SmoothSmem), (cudaStream_t)stream>>>(
      s->dev, n->dev, n->w_full, n->w_full_off, n->w_rows, n->w_rows_off, d_pts, d_flags, n_pts, d_grad, d_smooth,
      reinterpret_cast<float2*>(n->scratch));
  SURF_LAUNCH_CHECK();
  return 0;
}
