/* surf_b200 — C ABI of the B200-native SuRF volume-rendering hot path.
 *
 * One shared library (surf_b200/csrc/libsurf_b200.so, sm_100a only).  Plain
 * pointers and sizes; no torch / C++ types cross this boundary.  The reference
 * has no FFI on this path: its boundary is the Python module API of
 * models/modules/implicit_surface.py.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference root); the
 * Python mirror in surf_b200/modules/ binds them through ctypes
 * (INTEGRATION.md shows the binding a maintainer would add).
 *
 * Conventions
 *  - Every `const T* d_...` / `T* d_...` argument is DEVICE memory owned by the
 *    caller (PyTorch allocates it); `h_...` is HOST memory.  The library never
 *    frees or retains caller buffers.  The only library-owned device memory
 *    lives behind the opaque `surf_scene` / `surf_net` handles (explicit
 *    create/destroy).
 *  - All work is enqueued on the caller's `stream` (a cudaStream_t passed as
 *    void*); calls are asynchronous and re-entrant, no host sync inside unless
 *    stated.
 *  - Return code: 0 = ok, < 0 = invalid argument, > 0 = cudaError_t.  Nothing
 *    throws across the ABI; `surf_last_error()` returns a thread-local message.
 *  - There is NO CPU fallback: without a CUDA device every compute entry point
 *    returns an error.
 */
#ifndef SURF_B200_H
#define SURF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURF_ABI_VERSION 4
#define SURF_MAX_LEVELS 4
#define SURF_MAX_VIEWS 8       /* source views (nv-1) */
#define SURF_MAX_STAGES 4
#define SURF_SDF_LAYERS 7      /* lin0..lin6 (sdf_network.py:45-91, n_layers=6) */

typedef struct surf_scene surf_scene; /* opaque: prepared scene tensors */
typedef struct surf_net surf_net;     /* opaque: folded / re-laid-out network weights */

/* ---- scene ------------------------------------------------------------------------------
 * Inputs are the reference-layout tensors SuRF.build_volumes emits (surf.py:80-131,
 * volume.py:99-132) in RENDERER order (lists already reversed, surf.py:159):
 * volume lists fine->coarse, feature lists high-res->low-res. */
typedef struct surf_scene_inputs {
  int32_t n_levels;                             /* 1..4; 0 = matching-volume-only scene (surf_depth_map) */
  int32_t feat_ch;                              /* channels per level (7) */
  int32_t dim[SURF_MAX_LEVELS];                 /* cubic volume dim N_l */
  int64_t n_vox[SURF_MAX_LEVELS];               /* rows of volumes[l] */
  const float* d_volumes[SURF_MAX_LEVELS];      /* (n_vox_l, feat_ch) fp32            surf.py:119 */
  const int64_t* d_sparse_idx[SURF_MAX_LEVELS]; /* (N,N,N) int64, -1 = empty          volume.py:123-132 */
  const float* d_mask_volumes[SURF_MAX_LEVELS]; /* (1,1,N,N,N) fp32 0/1               volume.py:112-119 */
  const float* d_matching_volume;               /* (1,1,M,M,M) fp32 logits, may be NULL (no sampler) */
  int32_t match_dim;
  int32_t n_views;                              /* nv (reference view 0 + sources) */
  int32_t img_h, img_w;
  int32_t n_feat_levels;                        /* 4 */
  const float* d_imgs;                          /* (nv,3,H,W) fp32                    dtu.py:318,384 */
  const float* d_features[4];                   /* (nv,4,H>>i,W>>i) fp32 NCHW         feature_network.py:178 */
  const float* h_intrs;                         /* HOST (nv,4,4) row-major */
  const float* h_w2cs;                          /* HOST (nv,4,4) = inverse(c2ws), computed by the caller
                                                   with the same routine as the reference (torch.inverse,
                                                   projector.py:529) */
  const float* h_c2ws;                          /* HOST (nv,4,4) */
} surf_scene_inputs;

typedef struct surf_scene_stats {
  int64_t bytes_index, bytes_volumes, bytes_masks, bytes_matching, bytes_images;
  int64_t n_vox[SURF_MAX_LEVELS];
} surf_scene_stats;

/* scene_prepare: int64->int32 index tables, fp32->1-bit masks, 7->8-float voxel rows,
 * NCHW->NHWC feature maps (level 0 fused with RGB into 32-byte texels).  Replaces nothing in
 * the reference (it consumes the raw layouts directly); once per scene. */
int surf_scene_create(const surf_scene_inputs* in, void* stream, surf_scene** out);
void surf_scene_destroy(surf_scene* s);
int surf_scene_get_stats(const surf_scene* s, surf_scene_stats* out);
/* The per-batch part of a scene: source images, feature pyramids and cameras (the `ipts` keys imgs / intrs / c2ws
 * and the feature lists of surf.py:138-159).  surf_scene_create installs them when given; surf_scene_set_views
 * replaces them on a live scene without touching the (multi-GB) volume part, e.g. when SuRF.forward selects other
 * `view_ids` (surf.py:140-146).  d_imgs may be NULL (cameras only).  Stream-ordered. */
typedef struct surf_scene_views {
  int32_t n_views;                              /* nv */
  int32_t img_h, img_w;
  int32_t n_feat_levels;                        /* 4 */
  const float* d_imgs;                          /* (nv,3,H,W) */
  const float* d_features[4];                   /* (nv,4,H>>i,W>>i) NCHW */
  const float* h_intrs;                         /* HOST (nv,4,4) */
  const float* h_w2cs;                          /* HOST (nv,4,4) = inverse(c2ws) */
  const float* h_c2ws;                          /* HOST (nv,4,4) */
} surf_scene_views;
int surf_scene_set_views(surf_scene* s, const surf_scene_views* in, void* stream);
/* finetune mode optimises volumes[l] in place (surf.py:43-44,72): refresh the padded copy */
int surf_scene_update_volume(surf_scene* s, int32_t level, const float* d_volume, int64_t n_vox, void* stream);

/* ---- network ----------------------------------------------------------------------------
 * HOST pointers to an ImplicitSurface state_dict (names in SURVEY.md §5). */
typedef struct surf_net_inputs {
  /* SDFNetworkSparse (sdf_network.py:27-93) */
  int32_t n_lin;                           /* 7 */
  int32_t in_dim[SURF_SDF_LAYERS];         /* 27,156,156,156,156,156,156 */
  int32_t out_dim[SURF_SDF_LAYERS];        /* 128,128,101,128,128,128,129 */
  const float* h_weight_v[SURF_SDF_LAYERS];/* (out,in)  weight_v, or the plain weight if h_weight_g NULL */
  const float* h_weight_g[SURF_SDF_LAYERS];/* (out,)    nullable */
  const float* h_bias[SURF_SDF_LAYERS];    /* (out,) */
  int32_t multires;                        /* 4 -> PE dim 27 (embedder.py:39-51) */
  int32_t skip_layer;                      /* 3, or -1 */
  int32_t feat_channels;                   /* 28 */
  float scale;                             /* 1.0 */
  /* BlendingNetwork (blending_network.py:27-67): weight (out,in) + bias per Linear */
  const float* h_blend_w[11];              /* ray_dir_fc.0,.2 base_fc.0,.2 vis_fc.0,.2 vis_fc2.0,.2 rgb_fc.0,.2,.4 */
  const float* h_blend_b[11];
  float blend_s;                           /* anti-alias pooling scalar `s` */
  int32_t d_feature;                       /* 16 */
  /* SingleVarianceNetwork (variance_network.py:5-11) */
  float variance;
} surf_net_inputs;

int surf_net_create(const surf_net_inputs* in, void* stream, surf_net** out);
void surf_net_destroy(surf_net* n);

/* ---- MLP kernel family, chosen PER CALL (no process-wide state) ---------------------------------
 *   SURF_MLP_FFMA    fp32 CUDA-core kernels (sdf_mlp.cu, blend.cu): the parity anchor written first;
 *   SURF_MLP_TC      tcgen05 / TMEM kernels (sdf_tc2.cu, blend_tc.cu), every product as the fp16 hi/lo 3-MMA
 *                    split with fp32 accumulation: fp32-grade (1e-4) and bitwise reproducible.  This is what the
 *                    Python mirror passes by default and what bench.py measures;
 *   SURF_MLP_TC_FAST the same kernels with ONE fp16 MMA per product: the opt-in reduced-precision mode
 *                    (north_star: 1e-2 relative).
 * A network whose shape the tensor-core kernels do not support (multires != 4 or skip layer != 3) runs the
 * FFMA kernels whatever the mode. */
#define SURF_MLP_FFMA 0
#define SURF_MLP_TC 1
#define SURF_MLP_TC_FAST 4

/* ---- render configuration (confs/surf.conf: model.implicit_surface.render) ---------------- */
/* surf_render_cfg.color_path: how the projection gather and the blending network of a call are scheduled.  All three
 * give the same results bit for bit (tests/test_gpu_bench_scene.py); measured on the 576x800 image:
 * SERIAL 123.3 ms, OVERLAP 122.3 ms (the SDF kernel slows by what the gather gains), FUSED 126.6 ms. */
#define SURF_COLOR_SERIAL 0   /* gather kernel, then blending kernel, on the caller's stream (default) */
#define SURF_COLOR_OVERLAP 1  /* gather kernel on a library-owned side stream beside the SDF kernel, joined before the blend */
#define SURF_COLOR_FUSED 2    /* gather inside the blending kernel (tensor-core modes, 2 or 4 source views; else SERIAL) */

typedef struct surf_render_cfg {
  int32_t n_stages;                        /* 4 */
  int32_t n_samples[SURF_MAX_STAGES];      /* 64,32,24,16 */
  float sample_ranges[SURF_MAX_STAGES];    /* 1.0,0.4,0.1,0.01 */
  int32_t n_depth;                         /* 256 probe depths */
  int32_t perturb;                         /* jitter on/off (perturb > 0) */
  float cos_anneal_ratio;                  /* 1.0 in validation */
  int32_t chunk_rays;                      /* rays per reference render() call: the empty-mask fallback
                                              (implicit_surface.py:88-89) is evaluated per chunk; 0 = all rays */
  const float* d_lin_tables;               /* torch.linspace(0,1,n) for n = n_samples[0..], then n_depth,
                                              concatenated (host-generated: linspace is not reproducible by
                                              a device formula) */
  int32_t mlp_mode;                        /* SURF_MLP_* kernel family for the SDF / blending MLPs of THIS call */
  int32_t color_path;                      /* SURF_COLOR_*: how the projection gather and the blending network of this
                                              call are scheduled (same results bit for bit) */
} surf_render_cfg;

/* Outputs of render_core; any pointer may be NULL (not written).  Shapes use B rays, S samples,
 * P = B*S.  Matches the inference keys of the dict built at implicit_surface.py:247-266. */
typedef struct surf_render_outputs {
  float* d_color_fine;        /* (B,3) */
  float* d_render_depth;      /* (B,) */
  float* d_sdf_depth;         /* (B,1) */
  float* d_normal;            /* (B,3) rotated by inv(c2w0[:3,:3]) */
  float* d_val_normal;        /* (B,3) sum g*w*inside_sphere, unrotated (validate(), :380-382) */
  float* d_weights;           /* (B,S) */
  float* d_weight_sum;        /* (B,1) */
  float* d_weight_max;        /* (B,1) */
  uint8_t* d_valid_mask;      /* (B,1) bool */
  float* d_inside_sphere;     /* (B,S) */
  float* d_mid_inside_sphere; /* (B,1) */
  float* d_mid_z_vals;        /* (B,S) */
  float* d_gradients;         /* (B,S,3)  REQUIRED */
  float* d_sdf;               /* (P,1)    REQUIRED (tail of `sparse_sdf`) */
  float* d_gradient_error_sums; /* (2,) [sum relax_inside*err, sum relax_inside]; accumulated (zero it first) */
  /* stage outputs for parity tests (nullable) */
  uint8_t* d_point_flags;     /* (P,) bit0 voxel mask, bit1 computed */
  float* d_point_color;       /* (P,3) */
  uint8_t* d_point_views;     /* (P,) bitmask of valid source views */
  int32_t* d_prev_idx;        /* (B,) first zero-crossing index */
  float* d_alpha;             /* (B,S) */
  /* inputs of the training extras (nullable): the raw zero-crossing depth of every ray, before the validity mask
   * (implicit_surface.py:210), and the largest sample depth of the call (:219; zero it first, accumulated by max) */
  float* d_z_cross;           /* (B,) */
  float* d_z_max;             /* (1,) */
} surf_render_outputs;

size_t surf_render_workspace_bytes(int64_t n_rays, int32_t n_samples_total, int32_t n_src_views);

/* ImplicitSurface.render lines 270-311: coarse-to-fine z sampling.  d_t_rand (B,n_stages) raw
 * U[0,1) draws (the 0.5 shift is applied inside), NULL = no jitter.  Outputs sorted z_vals (B,S)
 * and (optional) the expected surface depth (B,). */
int surf_sample_rays(const surf_scene* s, const surf_render_cfg* cfg, const float* d_rays_o,
                     const float* d_rays_d, const float* d_near, const float* d_far, const float* d_t_rand,
                     int64_t n_rays, float* d_z_vals, float* d_surf_z, void* stream);

/* ImplicitSurface.render_core (implicit_surface.py:64-266) for inference keys.  d_z_vals (B,S). */
int surf_render_core(const surf_scene* s, const surf_net* n, const surf_render_cfg* cfg, const float* d_rays_o,
                     const float* d_rays_d, const float* d_z_vals, int64_t n_rays, int32_t n_samples_total,
                     const surf_render_outputs* out, void* d_workspace, size_t workspace_bytes, void* stream);

/* ImplicitSurface.render = sample_rays + render_core on one stream; z_vals kept in workspace. */
int surf_render_rays(const surf_scene* s, const surf_net* n, const surf_render_cfg* cfg, const float* d_rays_o,
                     const float* d_rays_d, const float* d_near, const float* d_far, const float* d_t_rand,
                     int64_t n_rays, const surf_render_outputs* out, void* d_workspace, size_t workspace_bytes,
                     void* stream);

/* SDFNetworkSparse.sdf / .gradient (sdf_network.py:123-141) on a flat point list.
 * d_sdf (n,) required; d_grad (n,3) nullable (forward only when NULL). */
int surf_sdf_points(const surf_scene* s, const surf_net* n, const float* d_pts, int64_t n_pts, float* d_sdf,
                    float* d_grad, int32_t mlp_mode, void* stream);

/* SDFNetworkSparse.forward (sdf_network.py:95-121): the full (n, d_out) output [sdf / scale, lin6 rows 1..d_out-1].
 * The extra outputs are dead on the render path (implicit_surface.py:95-97); plain fp32 kernel, API completeness. */
int surf_sdf_full(const surf_scene* s, const surf_net* n, const float* d_pts, int64_t n_pts, float* d_out,
                  int32_t d_out_dim, void* stream);

/* SDFNetworkSparse.gradient's second return value (sdf_network.py:143-150): smooth = d/dx sum_j (d sdf / d x_j) =
 * Hessian . (1,1,1), analytically (forward-mode tangent through the reverse pass) — training only (smooth_error,
 * implicit_surface.py:172).  mlp_mode: SURF_MLP_FFMA = plain fp32 kernel; SURF_MLP_TC = tcgen05 kernel, primal and
 * tangent stream as two M = 128 GEMMs per layer, fp16 hi/lo 3-MMA split (fp32-grade); SURF_MLP_TC_FAST = one MMA.  d_flags (nullable): per-point byte, bit 1 = evaluate, else write zeros
 * (the reference's masked-out default, :99).  d_grad (nullable): the first-order gradient from the same pass. */
int surf_sdf_smooth(const surf_scene* s, const surf_net* n, const float* d_pts, int64_t n_pts, const uint8_t* d_flags,
                    float* d_grad, float* d_smooth, int32_t mlp_mode, void* stream);

/* extract_geometry's SDF query (implicit_surface.py:337-351): u[x,y,z] = -sdf(xs[x],ys[y],zs[z]) on the
 * tensor-product grid of the three coordinate tables (host torch.linspace, uploaded).  Dense (Q16).
 * d_u is (nx,ny,nz) row-major.  sparsify != 0: opt-in fast mode, points whose 4-level voxel mask is 0
 * get `fill` instead of an MLP evaluation (NOT result-identical outside the mask). */
int surf_sdf_grid(const surf_scene* s, const surf_net* n, const float* d_xs, int32_t nx, const float* d_ys,
                  int32_t ny, const float* d_zs, int32_t nz, float* d_u, int32_t sparsify, float fill,
                  int32_t mlp_mode, void* stream);

/* ---- stage-isolated entry points (parity tests; same kernels/device functions) ------------ */
/* lookup_volume(pts, mask_volumes, 'nearest').any(-1)  (projector.py:392-420, implicit_surface.py:86) */
int surf_point_mask(const surf_scene* s, const float* d_pts, int64_t n_pts, uint8_t* d_mask, void* stream);
/* lookup_sparse_volume (projector.py:377-390) -> (n, feat_ch*n_levels) */
int surf_lookup_sparse(const surf_scene* s, const float* d_pts, int64_t n_pts, float* d_feats, void* stream);
/* lookup_feature (projector.py:501-556) -> feat_views (n,V,19), ray_diff (n,V,4), mask (n,V) u8 */
int surf_lookup_feature(const surf_scene* s, const float* d_pts, int64_t n_pts, float* d_feat_views,
                        float* d_ray_diff, uint8_t* d_mask, void* stream);
/* BlendingNetwork.forward (blending_network.py:69-117) on given inputs -> rgb (n,3) */
int surf_blend(const surf_net* n, const float* d_feat_views, const float* d_ray_diff, const uint8_t* d_mask,
               int64_t n_pts, int32_t n_src_views, float* d_rgb, int32_t mlp_mode, void* stream);
/* render_core lines 75-89: z_vals -> mid_z (B,S), flags (P,) with the per-chunk fallback applied */
int surf_point_flags(const surf_scene* s, const surf_render_cfg* cfg, const float* d_rays_o, const float* d_rays_d,
                     const float* d_z_vals, int64_t n_rays, int32_t n_samples_total, float* d_mid_z,
                     uint8_t* d_flags, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- misc ------------------------------------------------------------------------------- */
/* ---- training extras ------------------------------------------------------------------------
 * The training-only tail of render_core (implicit_surface.py:218-245) + surface_patch_warp2 / patch_homography
 * (projector.py:560-645): surface point pts_sdf0 = o + d * clamp(z_cross) (Q15), its unit normal in the reference
 * camera frame (third SDF-MLP pass), and the plane-homography warp of a patch_size^2 pixel patch of the 12-channel
 * feature maps [features[0] | up(features[1]) | up(features[2])] of the scene's current views.
 * The 3x3 camera matrices are formed on the HOST with the reference's own torch ops (inverse, matmul):
 *   R0t = c2w0[:3,:3]^T, t0 = -R0t c2w0[:3,3], K0 = intr0[:3,:3], K0inv = inverse(intr)[0,:3,:3],
 *   Ksrc[v] = intr[v+1][:3,:3], Rrel[v] = c2w[v+1][:3,:3]^T c2w0[:3,:3], RC[v] = c2w[v+1][:3,:3]^T (C0 - C[v+1]).
 * Outputs: d_ref_val (1,B,P,12), d_src_val (V,B,P,12) fp32 (P = patch_size^2), d_pts_sdf0 (B,3), d_normal_sdf0 (B,3,
 * nullable).  Takes the scene non-const: the 12-channel maps live behind the handle and are rebuilt after
 * surf_scene_set_views. */
typedef struct surf_extras_params {
  float R0t[9], t0[3], K0[9], K0inv[9];
  float Ksrc[SURF_MAX_VIEWS][9], Rrel[SURF_MAX_VIEWS][9], RC[SURF_MAX_VIEWS][3];
  int32_t n_src;
  int32_t patch_size;           /* 11 */
} surf_extras_params;
size_t surf_extras_workspace_bytes(int64_t n_rays);
int surf_render_extras(surf_scene* s, const surf_net* n, const surf_extras_params* p, const float* d_rays_o,
                       const float* d_rays_d, const float* d_z_cross, const float* d_z_max, int64_t n_rays,
                       float* d_pts_sdf0, float* d_normal_sdf0, float* d_ref_val, float* d_src_val, void* d_workspace,
                       size_t workspace_bytes, int32_t mlp_mode, void* stream);

/* ---- volume producers ----------------------------------------------------------------------
 * models/modules/volume.py:54-168 — the functions that build the scene tensors.  Cameras are HOST matrices formed
 * like the reference (h_w2cs = inverse(c2ws), h_intrs = the 4x4 intrinsics, (nv,4,4) row-major).
 *   surf_volume_back_proj    Volume.back_proj_multiscale (:54-97): d_feats[s] = the feature scales that are summed
 *                            (feats[stage_idx:], NCHW (nv,4,h_s,w_s)); norm_h/norm_w = size of the finest map (the
 *                            projection is normalised with it); h_agg_mlp = [0.weight (8,4), 0.bias (8), 2.weight (1,8),
 *                            2.bias (1)] of Volume.agg_mlp.  -> d_feat_vol (n,8) = [mean4 | var4], d_mask_vol (n) 0/1.
 *   surf_volume_depth_filter Volume.depth_filtering (:134-168): d_depths (nv,h,w), norm_h/w = (h,w); -> d_valid (n) 0/1.
 *   surf_volume_upsample2x   F.interpolate(scale_factor=2, "trilinear") of a (D,H,W) volume (sparse2dense, :108).
 *   surf_scene_create_sparse sparse2dense + get_index (:99-132) for all levels, emitting the prepared layout of the
 *                            render path directly (int32 index, 1-bit masks, 8-float voxel rows, fp32 matching volume
 *                            = finest logits scattered over the 2x up-sampled coarser volumes).  Levels in BUILD order
 *                            (coarse -> fine, dims doubling); d_coords[b] (n_b,3) fp32 integer voxel coordinates,
 *                            row i of d_volumes[b] (n_b,feat_ch) belongs to coordinate i; d_logits[b] (n_b) nullable
 *                            (no sampler).  The result is a regular scene handle; views are installed with surf_scene_set_views. */
typedef struct surf_volume_views {
  int32_t n_views;
  int32_t norm_h, norm_w;
  const float* h_w2cs;                          /* HOST (nv,4,4) = inverse(c2ws) */
  const float* h_intrs;                         /* HOST (nv,4,4) */
} surf_volume_views;
typedef struct surf_scene_sparse_inputs {
  int32_t n_levels;
  int32_t feat_ch;
  int32_t dim[SURF_MAX_LEVELS];
  int64_t n_vox[SURF_MAX_LEVELS];
  const float* d_coords[SURF_MAX_LEVELS];
  const float* d_volumes[SURF_MAX_LEVELS];
  const float* d_logits[SURF_MAX_LEVELS];
} surf_scene_sparse_inputs;
int surf_volume_back_proj(const surf_volume_views* vw, const float* const* d_feats, const int32_t* feat_h,
                          const int32_t* feat_w, int32_t n_scales, int32_t n_channels, const float* h_agg_mlp,
                          const float* d_coords, int64_t n_vox, const float* voxel_size, const float* origin,
                          float* d_feat_vol, uint8_t* d_mask_vol, void* stream);
int surf_volume_depth_filter(const surf_volume_views* vw, const float* d_depths, const float* d_coords, int64_t n_vox,
                             const float* voxel_size, const float* origin, float depth_range, uint8_t* d_valid,
                             void* stream);
int surf_volume_upsample2x(const float* d_in, int32_t D, int32_t H, int32_t W, float* d_out, void* stream);
int surf_scene_create_sparse(const surf_scene_sparse_inputs* in, void* stream, surf_scene** out);

/* ---- matching field -------------------------------------------------------------------------
 * MatchingField.forward for ONE view (matching_field.py:74-141): depth map of the view at (h,w) from the scene's dense
 * matching volume — per pixel 1 window [near, far] (stage 0) or 2 windows around the previous stage's depth (widths
 * range * ratio[0] and range * ratio[1], :101-121), n_samples uniform depths per window (+ jitter when d_t_rand (h*w,
 * n_windows) raw U[0,1) is given), sorted together, softmax of the trilinear probes -> expected depth * cos
 * (depth_render, :18-72) — then F.interpolate(bilinear) to (img_h, img_w) when d_depth_full is not NULL (:136).
 * Host-side (torch, like the reference): Kinv = intrs.inverse()[view,:3,:3], R = c2w[:3,:3], C = c2w[:3,3],
 * Rinv2 = row 2 of inverse(c2w[:3,:3]); d_lin = linspace(0,1,n_samples), d_tx = linspace(0,img_w-1,w),
 * d_ty = linspace(0,img_h-1,h); d_pre_depth (img_h,img_w) = previous stage's full-resolution map.
 * d_occ_partials (h*w,3): per ray [sum of the first 6 densities, sum density * outside_sphere, sum outside_sphere]
 * (occ_reg = col0.sum() / (6 h w) + col1.sum() / (col2.sum() + 1e-10), :68). */
typedef struct surf_depth_map_params {
  float Kinv[9], R[9], C[3], Rinv2[3];
  float near, far;
  float ratio[2];               /* range_ratios[stage], range_ratios[stage-1] */
  int32_t n_windows, n_samples;
  int32_t h, w, img_h, img_w;
} surf_depth_map_params;
int surf_depth_map(const surf_scene* s, const surf_depth_map_params* p, const float* d_lin, const float* d_tx,
                   const float* d_ty, const float* d_pre_depth, const float* d_t_rand, float* d_depth_lowres,
                   float* d_occ_partials, float* d_depth_full, void* stream);

/* ---- marching cubes -----------------------------------------------------------------------
 * Replaces the host call `mcubes.marching_cubes(u, threshold)` of extract_geometry (implicit_surface.py:353; PyMCubes
 * 0.1.4 is an un-vendored dependency of the reference).  u is the (nx,ny,nz) row-major fp32 grid surf_sdf_grid wrote;
 * a corner is inside when u > threshold (u = -sdf).  Two calls so that the caller owns every buffer:
 *   surf_mc_count  classifies the grid into d_workspace (surf_mc_workspace_bytes) and writes
 *                  d_counts[0] = number of vertices, d_counts[1] = number of triangles (device int64[2]);
 *   surf_mc_emit   (same u / dims / threshold / workspace) writes the vertices — (n_vertices,3) fp64 in grid-index
 *                  coordinates like PyMCubes, x shifted by x_offset for an x-slab — and the (n_triangles,3) int32 vertex
 *                  ids, oriented with the normal pointing outside (towards smaller u).
 * Every mesh vertex lies on a grid edge at the linear zero of u - threshold; vertices are shared between the
 * triangles of neighbouring cells (indexed mesh, watertight inside the grid). */
size_t surf_mc_workspace_bytes(int32_t nx, int32_t ny, int32_t nz);
int surf_mc_count(const float* d_u, int32_t nx, int32_t ny, int32_t nz, float threshold, void* d_workspace,
                  size_t workspace_bytes, int64_t* d_counts, void* stream);
int surf_mc_emit(const float* d_u, int32_t nx, int32_t ny, int32_t nz, float threshold, const void* d_workspace,
                 int32_t x_offset, double* d_vertices, int64_t n_vertices, int32_t* d_triangles, int64_t n_triangles,
                 void* stream);

/* ---- mesh cleaning after extract_geometry (utils/clean_mesh.py:10-129, runner.py:233 `--clean_mesh`) ----
 * Replaces the host pipeline skimage.binary_dilation -> torch projection -> trimesh/embree ray casting ->
 * trimesh connected components.  All buffers are the caller's; h_* pointers are HOST arrays (row-major).
 *   surf_mask_dilate              binary dilation of (n_views,h,w) byte masks with skimage's disk(radius), zero border
 *                                 (:119-123); d_workspace: n_views*h*w int32
 *   surf_mesh_vertex_visibility   d_count[i] = number of views in which vertex i projects inside the image and onto a set
 *                                 pixel of the mask (bilinear tap, align_corners=True, zero padding) (:12-28);
 *                                 h_w2c (n_views,3,4) = inverse(c2w)[:3], h_K (n_views,3,3)
 *   surf_mesh_first_hits          one view: d_face_hit[f] = 1 for every face that is the first hit of a camera ray
 *                                 through the (hs,ws) sample grid linspace(0,h-1,hs) x linspace(0,w-1,ws) whose
 *                                 nearest-upsampled mask pixel is set (:41-78; the reference casts them with embree);
 *                                 d_face_hit is OR-ed into (zero it before the first view).  d_stats (int32[2]):
 *                                 [0] += masked rays without a hit, [1] = faces with a footprint above 4096 samples
 *                                 (more than 65536 of them is an error the caller must check)
 *   surf_mesh_components          d_label[f] = component id (smallest face index of the component) of the graph that
 *                                 links two faces sharing an edge no third face shares (trimesh face_adjacency);
 *                                 d_keep[f] = 1 when the component has at least min_len faces (:99-102) */
int surf_mask_dilate(const uint8_t* d_masks, int32_t n_views, int32_t h, int32_t w, int32_t radius, int32_t* d_workspace,
                     uint8_t* d_out, void* stream);
int surf_mesh_vertex_visibility(const float* d_vertices, int64_t n_vertices, const float* h_w2c, const float* h_K,
                                int32_t n_views, const uint8_t* d_masks, int32_t h, int32_t w, int32_t* d_count,
                                void* stream);
size_t surf_mesh_raster_workspace_bytes(int32_t hs, int32_t ws);
int surf_mesh_first_hits(const float* d_vertices, const int32_t* d_faces, int64_t n_faces, const float* h_w2c,
                         const float* h_c2w, const float* h_K, const uint8_t* d_mask, int32_t h, int32_t w, int32_t hs,
                         int32_t ws, void* d_workspace, size_t workspace_bytes, uint8_t* d_face_hit, int32_t* d_stats,
                         void* stream);
size_t surf_mesh_components_workspace_bytes(int64_t n_faces);
int surf_mesh_components(const int32_t* d_faces, int64_t n_faces, int32_t min_len, void* d_workspace,
                         size_t workspace_bytes, int32_t* d_label, uint8_t* d_keep, void* stream);

/* ---- 2-D feature pyramid (FeatureNetwork, models/modules/feature_network.py:126-178) ----
 * NCHW fp32 tensors.  The InstanceNorm2d + ReLU that follows every convolution is not materialised: a convolution
 * leaves the (sum, sum of squares) of every thread block's raw outputs per (image, channel) in d_partials — fp64 pairs,
 * [n * c_out planes][blocks per image], blocks per image = surf_fpn_conv_blocks / surf_fpn_deconv_blocks of the OUTPUT
 * size — surf_fpn_finish_stats sums them in a fixed order into (mean, 1/sqrt(var + eps)) pairs, and the consumer passes
 * those as d_in_stats to have relu((x - mean) * rstd) applied while it loads (NULL: the input is a plain tensor).
 *   surf_fpn_conv3x3        nn.Conv2d(c_in, c_out, 3, stride, padding 1, bias=False); c_out in {4, 8, 16, 32, 64};
 *                           d_partials may be NULL (output layers)
 *   surf_fpn_deconv3x3s2    nn.ConvTranspose2d(c_in, c_out, 3, stride 2, padding 1, output_padding 1, bias=False):
 *                           (h, w) -> (2h, 2w); weight (c_in, c_out, 3, 3); c_out in {8, 16, 32}
 *   surf_fpn_norm_relu_add  d_out = relu(norm(a)) + relu(norm(b))  (decoder output, feature_network.py:166); b NULL:
 *                           d_out = relu(norm(a)) */
int32_t surf_fpn_conv_blocks(int32_t h_out, int32_t w_out);
int32_t surf_fpn_deconv_blocks(int32_t h_out, int32_t w_out);
int surf_fpn_conv3x3(const float* d_x, const float* d_in_stats, const float* d_weight, int32_t n, int32_t c_in, int32_t h,
                     int32_t w, int32_t c_out, int32_t stride, float* d_out, double* d_partials, void* stream);
int surf_fpn_deconv3x3s2(const float* d_x, const float* d_in_stats, const float* d_weight, int32_t n, int32_t c_in,
                         int32_t h, int32_t w, int32_t c_out, float* d_out, double* d_partials, void* stream);
int surf_fpn_finish_stats(const double* d_partials, int32_t n_planes, int32_t blocks_per_plane, int64_t pixels_per_plane,
                          float eps, float* d_stats, void* stream);
int surf_fpn_norm_relu_add(const float* d_a, const float* d_stats_a, const float* d_b, const float* d_stats_b,
                           int32_t n_planes, int64_t pixels_per_plane, float* d_out, void* stream);

/* Host helper (no device work): advance torch's CPU mt19937 state by n_draws 32-bit draws without producing them
 * (one float32 of torch.rand = one draw).  The pointers address the fields of the blob torch.get_rng_state() returns
 * (CPUGeneratorImplStateLegacy: `left` int32 at byte 8, `next` uint64 at 16, `state[624]` uint64 at 24).  Used to keep
 * the reference's sequential jitter stream (implicit_surface.py:276,305,174) when an image is ray-sharded over ranks. */
int surf_mt19937_skip(uint64_t* h_state624, int32_t* h_left, uint64_t* h_next, uint64_t n_draws);

int surf_version(void);
const char* surf_last_error(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
int64_t surf_launch_count(void);
/* Optional per-kernel device timing for bench.py's roofline: when enabled, every launch of the
 * kernel classes below is bracketed by cudaEvents on its own stream.  surf_timing_read synchronises
 * the recorded events, returns the summed duration (ms) and launch count per class since the last
 * read, and resets.  kind: 0 = SDF MLP (fwd+grad), 1 = SDF MLP (fwd only), 2 = projection gather,
 * 3 = blending MLP, 4 = sampler, 5 = point flags/mask, 6 = compositing. */
#define SURF_TIMING_KINDS 7
int surf_timing_enable(int32_t on);
int surf_timing_read(double* ms_out /*[SURF_TIMING_KINDS]*/, int64_t* launches_out /*[SURF_TIMING_KINDS]*/);

/* Diagnostic: one 128 x N x K GEMM through the tcgen05 building blocks of the tensor-core MLP kernels
 * (A operand in TMEM, B in shared memory, fp32 accumulate in TMEM).  D (128,N) = A (128,K) * B (N,K)^T,
 * all fp32 row-major device buffers; split != 0 uses the fp16 hi/lo 3-MMA scheme of the fp32-parity mode. */
int surf_tc_selftest(const float* d_A, const float* d_B, float* d_D, int32_t K, int32_t N, int32_t split,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SURF_B200_H */
