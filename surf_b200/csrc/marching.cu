// Marching cubes on the extract_geometry grid u[x,y,z] (implicit_surface.py:353, `mcubes.marching_cubes(u, thr)`;
// PyMCubes is an un-vendored third-party dependency of the reference — SURVEY.md §8c — so the mesh is pinned by the
// CPU restatement in oracle/mc_oracle.py and by size-independent properties: watertight, outward oriented, every
// vertex on a grid edge at the linear zero of u).
//
// HBM-bound integer work: two streaming passes over the grid, one 4-byte code word per grid point in between.
//   pass 1  k_mc_count : per grid point  — which of its 3 forward edges (+x,+y,+z) cross the iso-value (each crossing
//                        edge owns one mesh vertex) and the 8-bit case of the cell whose lowest corner it is.  Packs
//                        [case 8 | edge flags 3 | vertex rank in block 10 | triangle rank in block 11] into a word and
//                        writes the block totals.
//   scan    k_mc_scan_seg / _top : exclusive scan of the per-block totals (524 288 blocks at 512^3) in two levels.
//   pass 2  k_mc_emit  : vertices (fp64, index coordinates like PyMCubes) and triangles (vertex ids looked up through
//                        the code words of the edge-owning neighbour points).
// Vertex order = grid order (x-major), x/y/z edge of a point in that order; triangle order = cell order.  A corner is
// INSIDE when u > thr (u = -sdf: inside the object), triangles are oriented with the normal pointing outside.
#include "mc_tables.cuh"
#include "surf_internal.cuh"

#define MC_BLOCK 256

struct McGrid {
  int nx, ny, nz;
  int64_t n;          // nx*ny*nz
};

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  // exclusive scan over a 256-thread block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < MC_BLOCK / 32; ++w) {
    const uint32_t s = s_warp[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

__device__ __forceinline__ void mc_point_code(const float* __restrict__ u, const McGrid g, int64_t p, float thr,
                                              uint32_t* flags, uint32_t* cse) {
  const int z = (int)(p % g.nz);
  const int y = (int)((p / g.nz) % g.ny);
  const int x = (int)(p / ((int64_t)g.nz * g.ny));
  const int64_t sx = (int64_t)g.ny * g.nz, sy = g.nz;
  const bool hx = x + 1 < g.nx, hy = y + 1 < g.ny, hz = z + 1 < g.nz;
  const bool i0 = __ldg(u + p) > thr;
  bool ix = false, iy = false, iz = false;
  uint32_t f = 0;
  if (hx) { ix = __ldg(u + p + sx) > thr; f |= (ix != i0) ? 1u : 0u; }
  if (hy) { iy = __ldg(u + p + sy) > thr; f |= (iy != i0) ? 2u : 0u; }
  if (hz) { iz = __ldg(u + p + 1) > thr;  f |= (iz != i0) ? 4u : 0u; }
  uint32_t c = 0;
  if (hx && hy && hz) {
    // corner c = cx + 2 cy + 4 cz
    c = (i0 ? 1u : 0u) | (ix ? 2u : 0u) | (iy ? 4u : 0u) | (iz ? 16u : 0u);
    c |= (__ldg(u + p + sx + sy) > thr) ? 8u : 0u;
    c |= (__ldg(u + p + sx + 1) > thr) ? 32u : 0u;
    c |= (__ldg(u + p + sy + 1) > thr) ? 64u : 0u;
    c |= (__ldg(u + p + sx + sy + 1) > thr) ? 128u : 0u;
  }
  *flags = f;
  *cse = c;
}

__global__ void __launch_bounds__(MC_BLOCK)
k_mc_count(const float* __restrict__ u, const McGrid g, float thr, uint32_t* __restrict__ code,
           uint32_t* __restrict__ blk_v, uint32_t* __restrict__ blk_t) {
  __shared__ uint32_t s_warp[MC_BLOCK / 32];
  const int64_t p = (int64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
  uint32_t f = 0, c = 0;
  if (p < g.n) mc_point_code(u, g, p, thr, &f, &c);
  // vertex count (<= 3 per point, 768 per block) and triangle count (<= 5 per cell, 1280 per block) share one scan
  const uint32_t nv = __popc(f), nt = c_mc_ntri[c];
  uint32_t tot;
  const uint32_t rk = block_excl_scan(nv | (nt << 16), s_warp, &tot);
  const uint32_t rv = rk & 0xffffu, rt = rk >> 16;
  if (p < g.n) code[p] = c | (f << 8) | (rv << 11) | (rt << 21);
  if (threadIdx.x == 0) { blk_v[blockIdx.x] = tot & 0xffffu; blk_t[blockIdx.x] = tot >> 16; }
}

// Exclusive scan of the per-block totals in two levels: k_mc_scan_seg scans segments of 1024 block totals in place
// (one CTA per segment) and writes the segment totals; k_mc_scan_top scans those (one CTA: 512 segments at 512^3) and
// writes the grand totals.  A block's offset is blk[b] + seg[b >> 10].
#define MC_SEG 1024
__device__ __forceinline__ void scan1024(unsigned long long (&v)[2], unsigned long long (&excl)[2],
                                         unsigned long long (&tot)[2], unsigned long long (*s_w)[32]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc[2];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    inc[a] = v[a];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, inc[a], o);
      if (lane >= o) inc[a] += t;
    }
    if (lane == 31) s_w[a][warp] = inc[a];
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    unsigned long long pre = 0ull, t = 0ull;
    for (int w = 0; w < 32; ++w) {
      const unsigned long long x = s_w[a][w];
      if (w < warp) pre += x;
      t += x;
    }
    excl[a] = pre + inc[a] - v[a];
    tot[a] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(MC_SEG)
k_mc_scan_seg(uint32_t* __restrict__ blk_v, uint32_t* __restrict__ blk_t, int64_t n, unsigned long long* __restrict__ seg) {
  __shared__ unsigned long long s_w[2][32];
  const int64_t i = (int64_t)blockIdx.x * MC_SEG + threadIdx.x;
  unsigned long long v[2] = {0ull, 0ull}, ex[2], tot[2];
  if (i < n) { v[0] = blk_v[i]; v[1] = blk_t[i]; }
  scan1024(v, ex, tot, s_w);
  if (i < n) { blk_v[i] = (uint32_t)ex[0]; blk_t[i] = (uint32_t)ex[1]; }
  if (threadIdx.x == 0) { seg[2 * blockIdx.x] = tot[0]; seg[2 * blockIdx.x + 1] = tot[1]; }
}

__global__ void __launch_bounds__(MC_SEG)
k_mc_scan_top(unsigned long long* __restrict__ seg, int64_t n_seg, int64_t* __restrict__ totals) {
  __shared__ unsigned long long s_w[2][32];
  unsigned long long carry[2] = {0ull, 0ull};
  for (int64_t base = 0; base < n_seg; base += MC_SEG) {
    const int64_t i = base + threadIdx.x;
    unsigned long long v[2] = {0ull, 0ull}, ex[2], tot[2];
    if (i < n_seg) { v[0] = seg[2 * i]; v[1] = seg[2 * i + 1]; }
    scan1024(v, ex, tot, s_w);
    if (i < n_seg) { seg[2 * i] = carry[0] + ex[0]; seg[2 * i + 1] = carry[1] + ex[1]; }
    carry[0] += tot[0]; carry[1] += tot[1];
  }
  if (threadIdx.x == 0) { totals[0] = (int64_t)carry[0]; totals[1] = (int64_t)carry[1]; }
}

__global__ void __launch_bounds__(MC_BLOCK)
k_mc_emit(const float* __restrict__ u, const McGrid g, float thr, const uint32_t* __restrict__ code,
          const uint32_t* __restrict__ blk_v, const uint32_t* __restrict__ blk_t,
          const unsigned long long* __restrict__ seg, int x_offset,
          double* __restrict__ verts, int64_t max_verts, int32_t* __restrict__ tris, int64_t max_tris) {
  const int64_t p = (int64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
  if (p >= g.n) return;
  const uint32_t w = code[p];
  const uint32_t c = w & 255u, f = (w >> 8) & 7u;
  if ((c == 0u || c == 255u) && f == 0u) return;
  const int z = (int)(p % g.nz);
  const int y = (int)((p / g.nz) % g.ny);
  const int x = (int)(p / ((int64_t)g.nz * g.ny));
  const int64_t strides[3] = {(int64_t)g.ny * g.nz, (int64_t)g.nz, 1};
  if (f) {
    int64_t vid = (int64_t)blk_v[blockIdx.x] + (int64_t)seg[2 * (blockIdx.x >> 10)] + ((w >> 11) & 1023u);
    const double u0 = (double)__ldg(u + p);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!((f >> a) & 1u)) continue;
      const double u1 = (double)__ldg(u + p + strides[a]);
      const double t = ((double)thr - u0) / (u1 - u0);
      if (vid < max_verts) {
        verts[vid * 3 + 0] = (double)(x + x_offset) + (a == 0 ? t : 0.0);
        verts[vid * 3 + 1] = (double)y + (a == 1 ? t : 0.0);
        verts[vid * 3 + 2] = (double)z + (a == 2 ? t : 0.0);
      }
      ++vid;
    }
  }
  const int nt = c_mc_ntri[c];
  if (nt == 0) return;
  int64_t tid = (int64_t)blk_t[blockIdx.x] + (int64_t)seg[2 * (blockIdx.x >> 10) + 1] + ((w >> 21) & 2047u);
  // vertex id of each of the 12 cell edges, looked up lazily (a cell uses 3..12 of them)
  for (int k = 0; k < nt; ++k, ++tid) {
    int32_t ids[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int e = c_mc_tris[c][3 * k + j];
      const int axis = e >> 2;
      const int lo = c_mc_edge_corner[e][0];                       // lower corner of the edge
      const int64_t q = p + (lo & 1 ? strides[0] : 0) + (lo & 2 ? strides[1] : 0) + (lo & 4 ? strides[2] : 0);
      const uint32_t wq = code[q];
      const uint32_t fq = (wq >> 8) & 7u;
      const uint32_t rank = __popc(fq & ((1u << axis) - 1u));
      const int64_t qb = q / MC_BLOCK;
      ids[j] = (int32_t)((int64_t)blk_v[qb] + (int64_t)seg[2 * (qb >> 10)] + ((wq >> 11) & 1023u) + rank);
    }
    if (tid < max_tris) {
      tris[tid * 3 + 0] = ids[0];
      tris[tid * 3 + 1] = ids[1];
      tris[tid * 3 + 2] = ids[2];
    }
  }
}

static int mc_grid(int32_t nx, int32_t ny, int32_t nz, McGrid* g, int64_t* n_blocks) {
  SURF_CHECK_ARG(nx >= 1 && ny >= 1 && nz >= 1, "grid dims must be positive");
  g->nx = nx; g->ny = ny; g->nz = nz;
  g->n = (int64_t)nx * ny * nz;
  *n_blocks = (g->n + MC_BLOCK - 1) / MC_BLOCK;
  SURF_CHECK_ARG(*n_blocks < (int64_t)0x7fffffff, "grid too large for one launch");
  return 0;
}

extern "C" size_t surf_mc_workspace_bytes(int32_t nx, int32_t ny, int32_t nz) {
  const int64_t n = (int64_t)nx * ny * nz;
  const int64_t nb = (n + MC_BLOCK - 1) / MC_BLOCK;
  // code words + two block arrays + segment totals + grand totals, each 256-byte aligned
  auto al = [](int64_t b) { return (size_t)((b + 255) / 256 * 256); };
  const int64_t ns = (nb + MC_SEG - 1) / MC_SEG;
  return al(n * 4) + 2 * al(nb * 4) + al(ns * 16) + 256;
}

static void mc_carve(void* ws, const McGrid& g, int64_t nb, uint32_t** code, uint32_t** bv, uint32_t** bt,
                     unsigned long long** seg, int64_t** tot) {
  auto al = [](int64_t b) { return (size_t)((b + 255) / 256 * 256); };
  const int64_t ns = (nb + MC_SEG - 1) / MC_SEG;
  char* p = (char*)ws;
  *code = (uint32_t*)p; p += al(g.n * 4);
  *bv = (uint32_t*)p;   p += al(nb * 4);
  *bt = (uint32_t*)p;   p += al(nb * 4);
  *seg = (unsigned long long*)p; p += al(ns * 16);
  *tot = (int64_t*)p;
}

extern "C" int surf_mc_count(const float* d_u, int32_t nx, int32_t ny, int32_t nz, float threshold, void* d_workspace,
                             size_t workspace_bytes, int64_t* d_counts, void* stream) {
  McGrid g;
  int64_t nb;
  int rc = mc_grid(nx, ny, nz, &g, &nb);
  if (rc) return rc;
  SURF_CHECK_ARG(d_u && d_workspace && d_counts, "null pointer");
  SURF_CHECK_ARG(workspace_bytes >= surf_mc_workspace_bytes(nx, ny, nz), "workspace too small");
  uint32_t *code, *bv, *bt;
  unsigned long long* seg;
  int64_t* tot;
  mc_carve(d_workspace, g, nb, &code, &bv, &bt, &seg, &tot);
  cudaStream_t st = (cudaStream_t)stream;
  k_mc_count<<<(unsigned)nb, MC_BLOCK, 0, st>>>(d_u, g, threshold, code, bv, bt);
  SURF_LAUNCH_CHECK();
  const int64_t ns = (nb + MC_SEG - 1) / MC_SEG;
  k_mc_scan_seg<<<(unsigned)ns, MC_SEG, 0, st>>>(bv, bt, nb, seg);
  SURF_LAUNCH_CHECK();
  k_mc_scan_top<<<1, MC_SEG, 0, st>>>(seg, ns, tot);
  SURF_LAUNCH_CHECK();
  SURF_CUDA(cudaMemcpyAsync(d_counts, tot, 2 * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int surf_mc_emit(const float* d_u, int32_t nx, int32_t ny, int32_t nz, float threshold,
                            const void* d_workspace, int32_t x_offset, double* d_vertices, int64_t n_vertices,
                            int32_t* d_triangles, int64_t n_triangles, void* stream) {
  McGrid g;
  int64_t nb;
  int rc = mc_grid(nx, ny, nz, &g, &nb);
  if (rc) return rc;
  SURF_CHECK_ARG(d_u && d_workspace, "null pointer");
  SURF_CHECK_ARG(n_vertices < (int64_t)0x7fffffff, "more than 2^31 vertices");
  if (n_vertices == 0 && n_triangles == 0) return 0;
  SURF_CHECK_ARG(d_vertices && (d_triangles || n_triangles == 0), "null output");
  uint32_t *code, *bv, *bt;
  unsigned long long* seg;
  int64_t* tot;
  mc_carve((void*)d_workspace, g, nb, &code, &bv, &bt, &seg, &tot);
  k_mc_emit<<<(unsigned)nb, MC_BLOCK, 0, (cudaStream_t)stream>>>(d_u, g, threshold, code, bv, bt, seg, x_offset, d_vertices,
                                                                n_vertices, d_triangles, n_triangles);
  SURF_LAUNCH_CHECK();
  return 0;
}
