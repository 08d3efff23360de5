// scene_prepare: converts the reference-layout scene tensors into the compact
// HBM layout the kernels read: int32 index tables, 1-bit masks, 32-byte voxel rows, NHWC feature maps.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "surf_internal.cuh"

// ---------------------------------------------------------------------------------------------
// error / accounting
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void surf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void surf_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// SM count of the CURRENT device (cached per device ordinal; a process may drive several GPUs)
int surf_num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (device, kernel),
// thread-safe (launchers may be called from several host threads / for several devices).
#include <mutex>
#include <set>
#include <utility>
int surf_ensure_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  SURF_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  const auto key = std::make_pair(dev, func);
  if (done.count(key)) return 0;
  SURF_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert(key);
  return 0;
}

// ---- optional kernel timing -------------------------------------------------------------------
#include <mutex>
#include <vector>
struct TimedLaunch { int kind; cudaEvent_t a, b; };
static bool g_timing = false;
static std::mutex g_timing_mu;
static std::vector<TimedLaunch> g_timed;
static thread_local cudaEvent_t g_pending_begin = nullptr;

void surf_time_begin(int kind, cudaStream_t st) {
  if (!g_timing) return;
  cudaEvent_t a;
  if (cudaEventCreate(&a) != cudaSuccess) return;
  cudaEventRecord(a, st);
  g_pending_begin = a;
}
void surf_time_end(int kind, cudaStream_t st) {
  if (!g_timing || !g_pending_begin) return;
  cudaEvent_t b;
  if (cudaEventCreate(&b) != cudaSuccess) return;
  cudaEventRecord(b, st);
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timed.push_back({kind, g_pending_begin, b});
  g_pending_begin = nullptr;
}
extern "C" int surf_timing_enable(int32_t on) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timing = on != 0;
  return 0;
}
extern "C" int surf_timing_read(double* ms_out, int64_t* launches_out) {
  SURF_CHECK_ARG(ms_out && launches_out, "null pointer");
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (int k = 0; k < SURF_TIMING_KINDS; ++k) { ms_out[k] = 0.0; launches_out[k] = 0; }
  for (auto& t : g_timed) {
    float ms = 0.f;
    cudaEventSynchronize(t.b);
    if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess && t.kind >= 0 && t.kind < SURF_TIMING_KINDS) {
      ms_out[t.kind] += ms;
      launches_out[t.kind] += 1;
    }
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  g_timed.clear();
  return 0;
}

extern "C" int surf_version(void) { return SURF_ABI_VERSION; }
extern "C" const char* surf_last_error(void) { return g_err; }
extern "C" int64_t surf_launch_count(void) { return (int64_t)g_launches.load(); }

// ---------------------------------------------------------------------------------------------
// conversion kernels (HBM-bound streaming; 128-bit accesses where the layout allows)
// ---------------------------------------------------------------------------------------------
__global__ void k_index64_to_32(const longlong2* __restrict__ in, int2* __restrict__ out, size_t n_pairs,
                                const int64_t* in1, int32_t* out1, size_t n) {   // (same buffers: the odd last element)
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t p = i; p < n_pairs; p += stride) {
    longlong2 v = in[p];
    out[p] = make_int2((int)v.x, (int)v.y);
  }
  if (i == 0 && (n & 1)) out1[n - 1] = (int32_t)in1[n - 1];
}

// one warp packs 32 consecutive floats into one word with a ballot
__global__ void k_pack_mask(const float* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
  const size_t n_words = (n + 31) / 32;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (size_t w = warp; w < n_words; w += n_warps) {
    const size_t i = w * 32 + lane;
    const bool b = (i < n) && (in[i] > 0.f);
    const uint32_t word = __ballot_sync(0xffffffffu, b);
    if (lane == 0) out[w] = word;
  }
}

__global__ void k_pad_volume(const float* __restrict__ in, float4* __restrict__ out, size_t nvox, int ch) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < nvox; i += stride) {
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = (c < ch) ? in[i * ch + c] : 0.f;
    out[i * 2] = make_float4(v[0], v[1], v[2], v[3]);
    out[i * 2 + 1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// (nv,4,h,w) -> (nv,h,w) float4
__global__ void k_feat_nhwc(const float* __restrict__ in, float4* __restrict__ out, int nv, int h, int w) {
  const size_t hw = (size_t)h * w;
  const size_t n = (size_t)nv * hw;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const size_t v = i / hw, p = i % hw;
    const float* b = in + v * 4 * hw + p;
    out[i] = make_float4(b[0], b[hw], b[2 * hw], b[3 * hw]);
  }
}

// imgs (nv,3,H,W) + features[0] (nv,4,H,W) -> (nv,H,W) x 2 float4  [r,g,b,f0 | f1,f2,f3,0]
__global__ void k_img0_nhwc(const float* __restrict__ img, const float* __restrict__ f0, float4* __restrict__ out,
                            int nv, int h, int w) {
  const size_t hw = (size_t)h * w;
  const size_t n = (size_t)nv * hw;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const size_t v = i / hw, p = i % hw;
    const float* a = img + v * 3 * hw + p;
    const float* b = f0 + v * 4 * hw + p;
    out[i * 2] = make_float4(a[0], a[hw], a[2 * hw], b[0]);
    out[i * 2 + 1] = make_float4(b[hw], b[2 * hw], b[3 * hw], 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
static int grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = (size_t)surf_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int scene_alloc(surf_scene* s, void** p, size_t bytes) {
  if (s->n_owned >= 64) {
    surf_set_error("scene: too many allocations");
    return -2;
  }
  SURF_CUDA(cudaMalloc(p, bytes < 16 ? 16 : bytes));
  s->owned[s->n_owned++] = *p;
  return 0;
}

extern "C" void surf_scene_destroy(surf_scene* s) {
  if (!s) return;
  for (int i = 0; i < s->n_owned; ++i) cudaFree(s->owned[i]);
  for (int i = 0; i < 4; ++i)
    if (s->view_buf[i]) cudaFree(s->view_buf[i]);
  if (s->warp12) cudaFree(s->warp12);
  delete s;
}

extern "C" int surf_scene_get_stats(const surf_scene* s, surf_scene_stats* out) {
  SURF_CHECK_ARG(s && out, "scene/stats null");
  *out = s->stats;
  return 0;
}

static int pad_volume(surf_scene* s, int l, const float* d_volume, int64_t nvox, cudaStream_t st) {
  if (nvox > 0) {
    k_pad_volume<<<grid_for((size_t)nvox, 256), 256, 0, st>>>(d_volume, (float4*)s->dev.vol8[l], (size_t)nvox,
                                                               s->dev.feat_ch);
    SURF_LAUNCH_CHECK();
  }
  return 0;
}

// for volume.cu (surf_scene_create_sparse)
int scene_alloc_pub(surf_scene* s, void** p, size_t bytes) { return scene_alloc(s, p, bytes); }
int scene_pad_volume_pub(surf_scene* s, int level, const float* d_volume, int64_t n_vox, cudaStream_t st) {
  void* p = nullptr;
  int rc = scene_alloc(s, &p, (size_t)n_vox * 32);
  if (rc) return rc;
  s->dev.vol8[level] = (const float4*)p;
  s->stats.bytes_volumes += (size_t)n_vox * 32;
  return pad_volume(s, level, d_volume, n_vox, st);
}

extern "C" int surf_scene_update_volume(surf_scene* s, int32_t level, const float* d_volume, int64_t n_vox,
                                        void* stream) {
  SURF_CHECK_ARG(s && d_volume, "scene/volume null");
  SURF_CHECK_ARG(level >= 0 && level < s->dev.n_levels, "level out of range");
  SURF_CHECK_ARG(n_vox == s->nvox[level], "voxel count changed");
  return pad_volume(s, level, d_volume, n_vox, (cudaStream_t)stream);
}

// Per-batch part of a scene: source images / feature maps (NHWC re-layout) and cameras.  May be called any number
// of times on a live scene (SuRF.forward selects `view_ids` per step, surf.py:140-146): buffers are reused when the
// shape is unchanged, everything is stream-ordered on `stream`.
extern "C" int surf_scene_set_views(surf_scene* s, const surf_scene_views* in, void* stream) {
  SURF_CHECK_ARG(s && in, "scene/views null");
  SURF_CHECK_ARG(in->n_views >= 1 && in->n_views - 1 <= SURF_MAX_VIEWS, "n_views out of range");
  SURF_CHECK_ARG(in->h_intrs && in->h_w2cs && in->h_c2ws, "camera matrices missing");
  cudaStream_t st = (cudaStream_t)stream;
  DevScene& d = s->dev;
  d.nv = in->n_views;
  d.V = in->n_views - 1;
  s->views_version++;
  if (in->d_imgs) {
    SURF_CHECK_ARG(in->n_feat_levels == 4, "4 feature pyramid levels expected");
    SURF_CHECK_ARG(in->img_h >= 8 && in->img_w >= 8, "image too small");
    for (int i = 0; i < 4; ++i) SURF_CHECK_ARG(in->d_features[i] != nullptr, "feature level null");
    d.H = in->img_h;
    d.W = in->img_w;
    s->stats.bytes_images = 0;
    for (int i = 0; i < 4; ++i) {
      d.fh[i] = in->img_h >> i;
      d.fw[i] = in->img_w >> i;
      const size_t ni = (size_t)d.nv * d.fh[i] * d.fw[i];
      const size_t bytes = ni * (i == 0 ? 32 : 16);
      if (s->view_bytes[i] != bytes) {
        if (s->view_buf[i]) SURF_CUDA(cudaFreeAsync(s->view_buf[i], st));
        s->view_buf[i] = nullptr;
        s->view_bytes[i] = 0;
        SURF_CUDA(cudaMallocAsync(&s->view_buf[i], bytes, st));
        s->view_bytes[i] = bytes;
      }
      if (i == 0) {
        d.img0 = (const float4*)s->view_buf[0];
        k_img0_nhwc<<<grid_for(ni, 256), 256, 0, st>>>(in->d_imgs, in->d_features[0], (float4*)s->view_buf[0], d.nv, d.H, d.W);
      } else {
        d.feat[i] = (const float4*)s->view_buf[i];
        k_feat_nhwc<<<grid_for(ni, 256), 256, 0, st>>>(in->d_features[i], (float4*)s->view_buf[i], d.nv, d.fh[i], d.fw[i]);
      }
      SURF_LAUNCH_CHECK();
      s->stats.bytes_images += bytes;
    }
  }
  for (int v = 0; v < d.V; ++v) {
    const float* w = in->h_w2cs + (size_t)(v + 1) * 16;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) d.w2c[v][r * 4 + c] = w[r * 4 + c];
    const float* k = in->h_intrs + (size_t)(v + 1) * 16;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) d.K[v][r * 3 + c] = k[r * 4 + c];
    const float* c2w = in->h_c2ws + (size_t)(v + 1) * 16;
    for (int r = 0; r < 3; ++r) d.cen[v][r] = c2w[r * 4 + 3];
  }
  for (int r = 0; r < 3; ++r) d.refcen[r] = in->h_c2ws[r * 4 + 3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) d.rot0inv[r * 3 + c] = in->h_w2cs[r * 4 + c];
  return 0;
}

extern "C" int surf_scene_create(const surf_scene_inputs* in, void* stream, surf_scene** out) {
  SURF_CHECK_ARG(in && out, "inputs/out null");
  SURF_CHECK_ARG(in->n_levels >= 0 && in->n_levels <= SURF_MAX_LEVELS, "n_levels must be 0..4");
  SURF_CHECK_ARG(in->n_levels >= 1 || in->d_matching_volume != nullptr, "a scene without volume levels needs a matching volume");
  SURF_CHECK_ARG(in->feat_ch >= 1 && in->feat_ch <= 7, "feat_ch must be 1..7");
  SURF_CHECK_ARG(in->n_views >= 0 && in->n_views - 1 <= SURF_MAX_VIEWS, "n_views out of range");
  cudaStream_t st = (cudaStream_t)stream;
  {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
      surf_set_error("no CUDA device: surf_b200 has no CPU fallback");
      return e != cudaSuccess ? (int)e : -3;
    }
  }
  surf_scene* s = new surf_scene();
  memset(s, 0, sizeof(*s));
  DevScene& d = s->dev;
  d.n_levels = in->n_levels;
  d.feat_ch = in->feat_ch;
  int rc = 0;
#define SC_TRY(x)                \
  do {                           \
    rc = (x);                    \
    if (rc != 0) {               \
      surf_scene_destroy(s);     \
      return rc;                 \
    }                            \
  } while (0)

  for (int l = 0; l < in->n_levels; ++l) {
    const int N = in->dim[l];
    if (N < 2 || (size_t)N * N * N > 0x7fffffffull * 2) {
      surf_set_error("volume dim %d unsupported", N);
      surf_scene_destroy(s);
      return -1;
    }
    if (!(in->d_sparse_idx[l] && (in->d_volumes[l] || in->n_vox[l] == 0))) {
      surf_set_error("level %d: null tensor", l);
      surf_scene_destroy(s);
      return -1;
    }
    d.dim[l] = N;
    d.voxel[l] = 2.0f / (float)(N - 1);
    const size_t n3 = (size_t)N * N * N;
    void* p = nullptr;
    SC_TRY(scene_alloc(s, &p, n3 * sizeof(int32_t)));
    d.index[l] = (const int32_t*)p;
    k_index64_to_32<<<grid_for(n3 / 2 + 1, 256), 256, 0, st>>>((const longlong2*)in->d_sparse_idx[l], (int2*)p,
                                                               n3 / 2, in->d_sparse_idx[l], (int32_t*)p, n3);
    surf_count_launch();
    s->stats.bytes_index += n3 * sizeof(int32_t);

    const size_t n_words = (n3 + 31) / 32;
    SC_TRY(scene_alloc(s, &p, n_words * sizeof(uint32_t)));
    d.mask[l] = (const uint32_t*)p;
    if (in->d_mask_volumes[l]) {
      k_pack_mask<<<grid_for(n_words * 32, 256), 256, 0, st>>>(in->d_mask_volumes[l], (uint32_t*)p, n3);
      surf_count_launch();
    } else {   // SDF-only scene (sdf_network.sdf(x, volumes, indexes)): no masks, every lookup is false
      rc = (int)cudaMemsetAsync(p, 0, n_words * sizeof(uint32_t), st);
      if (rc != 0) {
        surf_set_error("mask memset failed");
        surf_scene_destroy(s);
        return rc;
      }
    }
    s->stats.bytes_masks += n_words * sizeof(uint32_t);

    s->nvox[l] = in->n_vox[l];
    s->stats.n_vox[l] = in->n_vox[l];
    SC_TRY(scene_alloc(s, &p, (size_t)in->n_vox[l] * 32));
    d.vol8[l] = (const float4*)p;
    SC_TRY(pad_volume(s, l, in->d_volumes[l], in->n_vox[l], st));
    s->stats.bytes_volumes += (size_t)in->n_vox[l] * 32;
  }

  d.matching = nullptr;
  d.mdim = 0;
  if (in->d_matching_volume) {
    // kept in the caller's buffer?  No: the handle must outlive the caller's tensor -> private copy.
    const size_t n3 = (size_t)in->match_dim * in->match_dim * in->match_dim;
    void* p = nullptr;
    SC_TRY(scene_alloc(s, &p, n3 * sizeof(float)));
    rc = (int)cudaMemcpyAsync(p, in->d_matching_volume, n3 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (rc != 0) {
      surf_set_error("matching volume copy failed");
      surf_scene_destroy(s);
      return rc;
    }
    d.matching = (const float*)p;
    d.mdim = in->match_dim;
    s->stats.bytes_matching = n3 * sizeof(float);
  }

  if (in->n_views > 0) {
    surf_scene_views v;
    memset(&v, 0, sizeof(v));
    v.n_views = in->n_views;
    v.img_h = in->img_h;
    v.img_w = in->img_w;
    v.n_feat_levels = in->n_feat_levels;
    v.d_imgs = in->d_imgs;
    for (int i = 0; i < 4; ++i) v.d_features[i] = in->d_features[i];
    v.h_intrs = in->h_intrs;
    v.h_w2cs = in->h_w2cs;
    v.h_c2ws = in->h_c2ws;
    SC_TRY(surf_scene_set_views(s, &v, stream));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    surf_set_error("scene_prepare kernel failed: %s", cudaGetErrorString(e));
    surf_scene_destroy(s);
    return (int)e;
  }
#undef SC_TRY
  *out = s;
  return 0;
}
