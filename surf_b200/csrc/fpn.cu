// 2-D feature pyramid of SuRF (models/modules/feature_network.py:126-178, FeatureNetwork): per stage two 3x3
// convolutions (stride 2 into a coarser stage) each followed by InstanceNorm2d + ReLU, a ConvTranspose2d(3, stride 2)
// + InstanceNorm + ReLU decoder with skip additions, and one bias-free 3x3 output convolution per stage.  The images are
// small-channel (3 -> 8 -> 16 -> 32 -> 64), so this is HBM / L2-bound streaming work on the CUDA cores, not a tensor-core
// GEMM: the InstanceNorm + ReLU of a layer is never materialised — its consumer applies (x - mean) * rstd and max(., 0)
// while it loads, and every convolution leaves per-block sums / sums of squares of its own raw output for the next one.
//   k_fpn_conv3x3<COUT, STRIDE>   thread = one output pixel, all COUT accumulators in registers; 16x16 output tile,
//                                 input tile with halo staged in shared memory 8 channels at a time, weights broadcast
//   k_fpn_deconv3x3s2<COUT>       ConvTranspose2d(k 3, stride 2, padding 1, output_padding 1) as a gather per output pixel
//   k_fpn_finish_stats            per-block (sum, sum of squares) partials, fp64, fixed order -> mean, 1 / sqrt(var + eps)
//   k_fpn_norm_relu_add           decoder output: relu(norm(deconv)) + relu(norm(encoder))
#include <math.h>
#include <stdint.h>

#include "surf_internal.cuh"

#define FPN_TILE 16
#define FPN_CK(stride) ((stride) == 2 ? 4 : 8)   // input channels per shared-memory pass (static smem <= 48 KB)

// x[n][c][y][x] with the producer's InstanceNorm + ReLU applied on the fly (stats == nullptr: plain tensor)
__device__ __forceinline__ float fpn_load(const float* __restrict__ x, const float2* __restrict__ stats, int n, int c, int C,
                                          int H, int W, int y, int xx) {
  if (y < 0 || y >= H || xx < 0 || xx >= W) return 0.f;          // zero padding of the normalised tensor
  const float v = x[(((size_t)n * C + c) * H + y) * W + xx];
  if (!stats) return v;
  const float2 s = stats[n * C + c];
  return fmaxf((v - s.x) * s.y, 0.f);
}

// (sum, sum of squares) of one block's raw outputs per channel -> partials[(n * COUT + co) * n_blocks + block]
// (fixed order: the statistics are bitwise reproducible; an atomicAdd per warp and channel was most of the kernel time)
template <int COUT>
__device__ __forceinline__ void fpn_block_partials(const float (&acc)[COUT], bool ok, int tid, int n, int block, int n_blocks,
                                                   float (*s_red)[COUT][2], double2* __restrict__ partials) {
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    float s = ok ? acc[co] : 0.f, q = s * s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) { s_red[warp][co][0] = s; s_red[warp][co][1] = q; }
  }
  __syncthreads();
  if (tid < COUT) {
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s += (double)s_red[w][tid][0]; q += (double)s_red[w][tid][1]; }
    partials[((size_t)n * COUT + tid) * n_blocks + block] = make_double2(s, q);
  }
}

template <int COUT, int STRIDE>
__global__ void __launch_bounds__(FPN_TILE * FPN_TILE)
k_fpn_conv3x3(const float* __restrict__ x, const float2* __restrict__ in_stats, const float* __restrict__ w, int N, int Cin,
              int H, int W, int Ho, int Wo, float* __restrict__ out, double2* __restrict__ partials) {
  constexpr int TI = FPN_TILE * STRIDE + 2;          // input tile edge incl. halo (stride 2: 34, the last column unused)
  constexpr int CK = FPN_CK(STRIDE);
  __shared__ float s_in[CK][TI][TI + 1];
  __shared__ __align__(16) float s_w[CK][9][COUT];
  __shared__ float s_red[8][COUT][2];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * FPN_TILE + tx;
  const int n = blockIdx.z;
  const int ox = blockIdx.x * FPN_TILE + tx, oy = blockIdx.y * FPN_TILE + ty;
  const int ix0 = blockIdx.x * FPN_TILE * STRIDE - 1, iy0 = blockIdx.y * FPN_TILE * STRIDE - 1;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += CK) {
    const int nc = min(CK, Cin - c0);
    for (int i = tid; i < nc * TI * TI; i += FPN_TILE * FPN_TILE) {
      const int ck = i / (TI * TI), rem = i - ck * TI * TI, yy = rem / TI, xx = rem - yy * TI;
      s_in[ck][yy][xx] = fpn_load(x, in_stats, n, c0 + ck, Cin, H, W, iy0 + yy, ix0 + xx);
    }
    // weights (COUT, Cin, 3, 3) -> [ck][tap][cout]
    for (int i = tid; i < nc * 9 * COUT; i += FPN_TILE * FPN_TILE) {
      const int co = i % COUT, t = (i / COUT) % 9, ck = i / (COUT * 9);
      s_w[ck][t][co] = w[((size_t)co * Cin + c0 + ck) * 9 + t];
    }
    __syncthreads();
    for (int ck = 0; ck < nc; ++ck) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float v = s_in[ck][ty * STRIDE + t / 3][tx * STRIDE + t % 3];
        const float4* w4 = reinterpret_cast<const float4*>(&s_w[ck][t][0]);      // one 128-bit broadcast per 4 outputs
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          const float4 ww = w4[c4];
          acc[4 * c4] = fmaf(v, ww.x, acc[4 * c4]);
          acc[4 * c4 + 1] = fmaf(v, ww.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(v, ww.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(v, ww.w, acc[4 * c4 + 3]);
        }
      }
    }
    __syncthreads();
  }
  const bool ok = ox < Wo && oy < Ho;
  if (ok) {
#pragma unroll
    for (int co = 0; co < COUT; ++co) out[(((size_t)n * COUT + co) * Ho + oy) * Wo + ox] = acc[co];
  }
  if (partials)
    fpn_block_partials<COUT>(acc, ok, tid, n, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, s_red, partials);
}

// ConvTranspose2d(Cin, COUT, 3, stride 2, padding 1, output_padding 1): out (2H, 2W); weight (Cin, COUT, 3, 3)
template <int COUT>
__global__ void __launch_bounds__(256)
k_fpn_deconv3x3s2(const float* __restrict__ x, const float2* __restrict__ in_stats, const float* __restrict__ w, int N, int Cin,
                  int H, int W, float* __restrict__ out, double2* __restrict__ partials) {
  extern __shared__ __align__(16) float s_wd[];          // [cin][tap][cout]
  __shared__ float s_red[8][COUT][2];
  const int Ho = 2 * H, Wo = 2 * W;
  for (int i = threadIdx.x; i < Cin * 9 * COUT; i += blockDim.x) {
    const int co = i % COUT, t = (i / COUT) % 9, ci = i / (COUT * 9);
    s_wd[i] = w[((size_t)ci * COUT + co) * 9 + t];
  }
  __syncthreads();
  const int n = blockIdx.z;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  const bool ok = ox < Wo && oy < Ho;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
  if (ok) {
    for (int ky = 0; ky < 3; ++ky) {
      const int iy2 = oy + 1 - ky;
      if (iy2 < 0 || (iy2 & 1) || (iy2 >> 1) >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix2 = ox + 1 - kx;
        if (ix2 < 0 || (ix2 & 1) || (ix2 >> 1) >= W) continue;
        const int t = ky * 3 + kx;
        for (int ci = 0; ci < Cin; ++ci) {
          const float v = fpn_load(x, in_stats, n, ci, Cin, H, W, iy2 >> 1, ix2 >> 1);
          const float4* w4 = reinterpret_cast<const float4*>(s_wd + ((size_t)ci * 9 + t) * COUT);
#pragma unroll
          for (int c4 = 0; c4 < COUT / 4; ++c4) {
            const float4 ww = w4[c4];
            acc[4 * c4] = fmaf(v, ww.x, acc[4 * c4]);
            acc[4 * c4 + 1] = fmaf(v, ww.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(v, ww.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(v, ww.w, acc[4 * c4 + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int co = 0; co < COUT; ++co) out[(((size_t)n * COUT + co) * Ho + oy) * Wo + ox] = acc[co];
  }
  fpn_block_partials<COUT>(acc, ok, threadIdx.x, n, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, s_red, partials);
}

__global__ void k_fpn_finish_stats(const double2* __restrict__ partials, int n_planes, int n_blocks, double count, float eps,
                                   float2* __restrict__ stats) {
  const int plane = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (plane >= n_planes) return;
  double s = 0.0, q = 0.0;
  for (int b = lane; b < n_blocks; b += 32) {
    const double2 p = partials[(size_t)plane * n_blocks + b];
    s += p.x;
    q += p.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double mean = s / count;
    double var = q / count - mean * mean;        // biased, like InstanceNorm2d
    if (var < 0.0) var = 0.0;
    stats[plane] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

// out = relu(norm(a)) + relu(norm(b)); planes of `hw` elements
__global__ void k_fpn_norm_relu_add(const float* __restrict__ a, const float2* __restrict__ sa, const float* __restrict__ b,
                                    const float2* __restrict__ sb, int64_t n, int hw, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / hw);
    const float2 x = sa[p], y = sb[p];
    out[i] = fmaxf((a[i] - x.x) * x.y, 0.f) + fmaxf((b[i] - y.x) * y.y, 0.f);
  }
}
__global__ void k_fpn_norm_relu(const float* __restrict__ a, const float2* __restrict__ sa, int64_t n, int hw,
                                float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 x = sa[(int)(i / hw)];
    out[i] = fmaxf((a[i] - x.x) * x.y, 0.f);
  }
}

static inline int conv_out(int x, int stride) { return (x + 2 - 3) / stride + 1; }

template <int COUT>
static int conv_launch(const float* x, const float2* st_in, const float* w, int N, int Cin, int H, int W, int stride,
                       float* out, double2* partials, cudaStream_t st) {
  const int Ho = conv_out(H, stride), Wo = conv_out(W, stride);
  dim3 grid((Wo + FPN_TILE - 1) / FPN_TILE, (Ho + FPN_TILE - 1) / FPN_TILE, N), block(FPN_TILE, FPN_TILE);
  if (stride == 1) k_fpn_conv3x3<COUT, 1><<<grid, block, 0, st>>>(x, st_in, w, N, Cin, H, W, Ho, Wo, out, partials);
  else k_fpn_conv3x3<COUT, 2><<<grid, block, 0, st>>>(x, st_in, w, N, Cin, H, W, Ho, Wo, out, partials);
  SURF_LAUNCH_CHECK();
  return 0;
}

// thread blocks per image of a convolution with an (h_out, w_out) output: the length of a plane's partials row
extern "C" int32_t surf_fpn_conv_blocks(int32_t h_out, int32_t w_out) {
  return ((w_out + FPN_TILE - 1) / FPN_TILE) * ((h_out + FPN_TILE - 1) / FPN_TILE);
}
extern "C" int32_t surf_fpn_deconv_blocks(int32_t h_out, int32_t w_out) { return ((w_out + 31) / 32) * ((h_out + 7) / 8); }

extern "C" int surf_fpn_conv3x3(const float* d_x, const float* d_in_stats, const float* d_weight, int32_t n, int32_t c_in,
                                int32_t h, int32_t w, int32_t c_out, int32_t stride, float* d_out, double* d_partials,
                                void* stream) {
  SURF_CHECK_ARG(d_x && d_weight && d_out, "null pointer");
  SURF_CHECK_ARG(n >= 1 && c_in >= 1 && h >= 1 && w >= 1 && (stride == 1 || stride == 2), "shape / stride");
  cudaStream_t st = (cudaStream_t)stream;
  const float2* si = reinterpret_cast<const float2*>(d_in_stats);
  double2* pp = reinterpret_cast<double2*>(d_partials);
  switch (c_out) {
    case 4: return conv_launch<4>(d_x, si, d_weight, n, c_in, h, w, stride, d_out, pp, st);
    case 8: return conv_launch<8>(d_x, si, d_weight, n, c_in, h, w, stride, d_out, pp, st);
    case 16: return conv_launch<16>(d_x, si, d_weight, n, c_in, h, w, stride, d_out, pp, st);
    case 32: return conv_launch<32>(d_x, si, d_weight, n, c_in, h, w, stride, d_out, pp, st);
    case 64: return conv_launch<64>(d_x, si, d_weight, n, c_in, h, w, stride, d_out, pp, st);
    default: surf_set_error("fpn conv: c_out %d unsupported (4, 8, 16, 32, 64)", c_out); return -1;
  }
}

extern "C" int surf_fpn_deconv3x3s2(const float* d_x, const float* d_in_stats, const float* d_weight, int32_t n,
                                    int32_t c_in, int32_t h, int32_t w, int32_t c_out, float* d_out, double* d_partials,
                                    void* stream) {
  SURF_CHECK_ARG(d_x && d_weight && d_out && d_partials, "null pointer");
  SURF_CHECK_ARG(n >= 1 && c_in >= 1 && c_in <= 128 && h >= 1 && w >= 1, "shape");
  cudaStream_t st = (cudaStream_t)stream;
  const float2* si = reinterpret_cast<const float2*>(d_in_stats);
  double2* pp = reinterpret_cast<double2*>(d_partials);
  dim3 grid((2 * w + 31) / 32, (2 * h + 7) / 8, n);
  const size_t smem = (size_t)c_in * 9 * c_out * sizeof(float);
  SURF_CHECK_ARG(smem <= 200 * 1024, "fpn deconv: weights do not fit shared memory");
  {
    const void* fn = c_out == 8 ? (const void*)k_fpn_deconv3x3s2<8> : c_out == 16 ? (const void*)k_fpn_deconv3x3s2<16>
                                                                                 : (const void*)k_fpn_deconv3x3s2<32>;
    const int rc = surf_ensure_dyn_smem(fn, 200 * 1024);
    if (rc) return rc;
  }
  switch (c_out) {
    case 8: k_fpn_deconv3x3s2<8><<<grid, 256, smem, st>>>(d_x, si, d_weight, n, c_in, h, w, d_out, pp); break;
    case 16: k_fpn_deconv3x3s2<16><<<grid, 256, smem, st>>>(d_x, si, d_weight, n, c_in, h, w, d_out, pp); break;
    case 32: k_fpn_deconv3x3s2<32><<<grid, 256, smem, st>>>(d_x, si, d_weight, n, c_in, h, w, d_out, pp); break;
    default: surf_set_error("fpn deconv: c_out %d unsupported (8, 16, 32)", c_out); return -1;
  }
  SURF_LAUNCH_CHECK();
  return 0;
}

extern "C" int surf_fpn_finish_stats(const double* d_partials, int32_t n_planes, int32_t blocks_per_plane,
                                     int64_t pixels_per_plane, float eps, float* d_stats, void* stream) {
  SURF_CHECK_ARG(d_partials && d_stats && n_planes >= 1 && blocks_per_plane >= 1 && pixels_per_plane >= 1, "arguments");
  k_fpn_finish_stats<<<(n_planes + 3) / 4, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double2*>(d_partials),
                                                                          n_planes, blocks_per_plane,
                                                                          (double)pixels_per_plane, eps,
                                                                          reinterpret_cast<float2*>(d_stats));
  SURF_LAUNCH_CHECK();
  return 0;
}

extern "C" int surf_fpn_norm_relu_add(const float* d_a, const float* d_stats_a, const float* d_b, const float* d_stats_b,
                                      int32_t n_planes, int64_t pixels_per_plane, float* d_out, void* stream) {
  SURF_CHECK_ARG(d_a && d_stats_a && d_out && n_planes >= 1 && pixels_per_plane >= 1, "arguments");
  SURF_CHECK_ARG((d_b == nullptr) == (d_stats_b == nullptr), "b and its statistics go together");
  SURF_CHECK_ARG(pixels_per_plane < 0x7fffffffll, "plane too large");
  const int64_t n = (int64_t)n_planes * pixels_per_plane;
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)surf_num_sms() * 16;
  if (g > cap) g = cap;
  if (d_b)
    k_fpn_norm_relu_add<<<(int)g, 256, 0, (cudaStream_t)stream>>>(d_a, reinterpret_cast<const float2*>(d_stats_a), d_b,
                                                                  reinterpret_cast<const float2*>(d_stats_b), n,
                                                                  (int)pixels_per_plane, d_out);
  else
    k_fpn_norm_relu<<<(int)g, 256, 0, (cudaStream_t)stream>>>(d_a, reinterpret_cast<const float2*>(d_stats_a), n,
                                                              (int)pixels_per_plane, d_out);
  SURF_LAUNCH_CHECK();
  return 0;
}
