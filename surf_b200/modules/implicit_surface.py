"""ImplicitSurface — drop-in for models/modules/implicit_surface.py:50-436 (the render hot path).

Same constructor (a conf tree with ``render.*``, ``sdf_network``, ``color_network``,
``variance_network``), same method signatures, same output keys / dtypes / devices for the
inference outputs, same state_dict names.  All compute runs in csrc/*.cu through the C-ABI of
include/surf_b200.h; this file only draws the reference's random numbers on the host (quirk Q1),
builds linspace tables on the host (not reproducible by a device formula), allocates outputs and
enqueues the kernels.  There is no PyTorch / CPU fallback.

``render`` / ``render_core`` return all 18 keys of the reference, including the training extras ``smooth_error``
(analytic Hessian-vector kernel), ``ref_gray_val`` / ``sampled_gray_val`` (surface_patch_warp2); ``validate`` skips
them like it ignores them in the reference.  Not provided: autograd THROUGH the render (the backward pass of training /
finetuning) — outputs are detached.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, mesh
from ..scene import GLOBAL_SCENE_CACHE, PreparedScene
from .blending_network import BlendingNetwork
from .sdf_network import SDFNetworkSparse
from .variance_network import SingleVarianceNetwork

N_RANDOM_PTS = 1024     # implicit_surface.py:174
PATCH_SIZE = 11         # projector.py:560


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _NetHandle:
    def __init__(self, handle, lib):
        self.h, self.lib = handle, lib

    def __del__(self):
        try:
            if self.h is not None:
                self.lib.surf_net_destroy(self.h)
                self.h = None
        except Exception:
            pass


class ImplicitSurface(nn.Module):
    def __init__(self, confs):
        super().__init__()
        self.n_samples = [int(x) for x in confs.get_list("render.n_samples")]
        self.sample_ranges = [float(x) for x in confs.get_list("render.sample_ranges")]
        self.n_depth = confs.get_int("render.n_depth")
        self.perturb = confs.get_float("render.perturb")
        self.sdf_network = SDFNetworkSparse(**confs["sdf_network"])
        self.color_network = BlendingNetwork(**confs["color_network"])
        self.deviation_network = SingleVarianceNetwork(**confs["variance_network"])
        ref = weakref.ref(self)
        self.sdf_network._net_owner = ref
        self.color_network._net_owner = ref
        self._net = None
        self._net_key = None
        self._lin = None
        self._ws = None
        self.scene_cache = GLOBAL_SCENE_CACHE
        self.ray_batch = 1 << 15        # rays per launch set in validate(): 24-32 k is the measured optimum (tools/ray_batch_sweep.py)
        # MLP kernel family of every call made through this module (include/surf_b200.h SURF_MLP_*): the tcgen05
        # fp32-grade kernels by default; optional conf key ``mlp_mode`` (not a reference key) or assign the attribute
        mode = _lib.MLP_TC
        try:
            mode = confs.get_int("mlp_mode", default=_lib.MLP_TC)
        except Exception:
            pass
        self.mlp_mode = int(mode)

    # -- device handles --------------------------------------------------------------------------
    def net_handle(self):
        """surf_net* for the current parameter values (rebuilt when any parameter changed)."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._net is None or key != self._net_key:
            lib = _lib.load()
            inp = _lib.NetInputs()
            keep = []
            self.sdf_network.fill_net_inputs(inp, keep)
            self.color_network.fill_net_inputs(inp, keep)
            inp.variance = float(self.deviation_network.variance.detach().cpu())
            h = C.c_void_p()
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("surf_b200: move the module to a CUDA device first (no CPU fallback)")
            with torch.cuda.device(dev):
                _lib.check(lib.surf_net_create(C.byref(inp), _stream(), C.byref(h)), "net_create")
            self._net = _NetHandle(h, lib)
            self._net_key = key
        return self._net.h

    def _lin_tables(self, device):
        """torch.linspace(0,1,n) tables, generated on the host exactly as the reference does
        (implicit_surface.py:271,283,301 create them on the CPU then .type_as)."""
        if self._lin is None or self._lin.device != device:
            t = [torch.linspace(0.0, 1.0, n) for n in self.n_samples] + [torch.linspace(0.0, 1.0, self.n_depth)]
            self._lin = torch.cat(t).to(device)
        return self._lin

    def _cfg(self, device, cos_anneal_ratio, chunk_rays):
        cfg = _lib.RenderCfg()
        cfg.n_stages = len(self.n_samples)
        for i, n in enumerate(self.n_samples):
            cfg.n_samples[i] = n
            cfg.sample_ranges[i] = self.sample_ranges[i]
        cfg.n_depth = self.n_depth
        cfg.perturb = 1 if self.perturb > 0 else 0
        cfg.cos_anneal_ratio = float(cos_anneal_ratio)
        cfg.chunk_rays = int(chunk_rays)
        cfg.d_lin_tables = self._lin_tables(device).data_ptr()
        if self.mlp_mode not in _lib.MLP_MODES:
            raise ValueError("mlp_mode must be one of %s" % (_lib.MLP_MODES,))
        cfg.mlp_mode = int(self.mlp_mode)
        cfg.color_path = int(getattr(self, "color_path", _lib.COLOR_SERIAL))
        return cfg

    def _workspace(self, device, B, S, V):
        need = int(_lib.load().surf_render_workspace_bytes(B, S, V))
        if self._ws is None or self._ws.device != device or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def prepare(self, matching_volume, volumes, sparse_idxes, mask_volumes, imgs, features, intrs, c2ws):
        return self.scene_cache.get(volumes, sparse_idxes, mask_volumes, matching_volume, imgs, features, intrs, c2ws)

    @staticmethod
    def draw_t_rand(batch_size, n_stages=4):
        """Per-stage jitter draws from torch's global CPU generator in the reference's order (Q1):
        rand([B,1]) for stage 0 (implicit_surface.py:276) then one per fine stage (:305)."""
        return torch.cat([torch.rand([batch_size, 1]) for _ in range(n_stages)], dim=1)

    # -- kernels ---------------------------------------------------------------------------------
    def _render_device(self, scene, rays_o, rays_d, near, far, t_rand, z_vals, cos_anneal_ratio, chunk_rays,
                       stages=False, lean=False):
        lib = _lib.load()
        dev = rays_o.device
        B = rays_o.shape[0]
        S = sum(self.n_samples)
        f32 = dict(dtype=torch.float32, device=dev)
        rays_o = rays_o.detach().to(torch.float32).contiguous()
        rays_d = rays_d.detach().to(torch.float32).contiguous()
        cfg = self._cfg(dev, cos_anneal_ratio, chunk_rays)
        o = _lib.RenderOutputs()
        t = {}

        def out(name, shape, dtype=torch.float32, zero=False):
            x = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=dev)
            t[name] = x
            setattr(o, "d_" + name, x.data_ptr())
            return x

        out("color_fine", (B, 3))
        out("render_depth", (B,))
        out("sdf_depth", (B, 1))
        out("val_normal", (B, 3))
        out("gradients", (B, S, 3))
        out("sdf", (B * S, 1))
        if not lean:
            out("normal", (B, 3))
            out("weights", (B, S))
            out("weight_sum", (B, 1))
            out("weight_max", (B, 1))
            out("valid_mask", (B, 1), torch.uint8)
            out("inside_sphere", (B, S))
            out("mid_inside_sphere", (B, 1))
            out("mid_z_vals", (B, S))
            out("gradient_error_sums", (2,), zero=True)
            out("z_cross", (B,))
            out("z_max", (1,), zero=True)
        if not lean and not stages:
            out("point_flags", (B * S,), torch.uint8)       # the second-order extra needs the evaluated-point bits
        if stages:
            out("point_flags", (B * S,), torch.uint8)
            out("point_color", (B * S, 3), zero=True)
            out("point_views", (B * S,), torch.uint8, zero=True)
            out("prev_idx", (B,), torch.int32)
            out("alpha", (B, S))
        ws = self._workspace(dev, B, S, scene.n_src_views)
        net = self.net_handle()
        if z_vals is None:
            near = near.detach().to(torch.float32).contiguous()
            far = far.detach().to(torch.float32).contiguous()
            tr = None
            if t_rand is not None and self.perturb > 0:
                tr = t_rand.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
            with torch.cuda.device(dev):
                _lib.check(lib.surf_render_rays(scene.handle, net, C.byref(cfg), rays_o.data_ptr(), rays_d.data_ptr(),
                                                near.data_ptr(), far.data_ptr(),
                                                tr.data_ptr() if tr is not None else None, B, C.byref(o), ws.data_ptr(),
                                                ws.numel(), _stream()), "render_rays")
        else:
            z = z_vals.detach().to(torch.float32).contiguous()
            with torch.cuda.device(dev):
                _lib.check(lib.surf_render_core(scene.handle, net, C.byref(cfg), rays_o.data_ptr(), rays_d.data_ptr(),
                                                z.data_ptr(), B, S, C.byref(o), ws.data_ptr(), ws.numel(), _stream()),
                           "render_core")
        return t

    def sample_z(self, scene, rays_o, rays_d, near, far, t_rand):
        """render() lines 270-311 only: sorted z_vals (B,S) and the expected surface depth (B,)."""
        lib = _lib.load()
        dev = rays_o.device
        B, S = rays_o.shape[0], sum(self.n_samples)
        cfg = self._cfg(dev, 1.0, 0)
        z = torch.empty((B, S), dtype=torch.float32, device=dev)
        surf = torch.empty((B,), dtype=torch.float32, device=dev)
        rays_o = rays_o.detach().float().contiguous()
        rays_d = rays_d.detach().float().contiguous()
        near = near.detach().float().contiguous()
        far = far.detach().float().contiguous()
        tr = t_rand.to(device=dev, dtype=torch.float32).contiguous() if (t_rand is not None and self.perturb > 0) else None
        with torch.cuda.device(dev):
            _lib.check(lib.surf_sample_rays(scene.handle, C.byref(cfg), rays_o.data_ptr(), rays_d.data_ptr(),
                                            near.data_ptr(), far.data_ptr(), tr.data_ptr() if tr is not None else None,
                                            B, z.data_ptr(), surf.data_ptr(), _stream()), "sample_rays")
        return z, surf

    def _sparse_sdf_random(self, scene, pts_random, device):
        """implicit_surface.py:174-178: SDF of the random points that fall inside the voxel mask."""
        lib = _lib.load()
        p = pts_random.to(device=device, dtype=torch.float32).contiguous()
        m = torch.empty(p.shape[0], dtype=torch.uint8, device=device)
        s = torch.empty((p.shape[0], 1), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.check(lib.surf_point_mask(scene.handle, p.data_ptr(), p.shape[0], m.data_ptr(), _stream()),
                       "point_mask")
            _lib.check(lib.surf_sdf_points(scene.handle, self.net_handle(), p.data_ptr(), p.shape[0], s.data_ptr(),
                                           None, int(self.mlp_mode), _stream()), "sdf_points")
        # the reference scatters the masked points into zeros (:175-178): exactly 0 outside the mask, also where the
        # extrapolating trilinear weights of a far-away point overflow (inf * 0 would be NaN)
        return torch.where(m[:, None].bool(), s, torch.zeros_like(s))

    def _render_extras(self, t, scene, rays_o, rays_d, intrs, c2ws):
        """The training-only tail of render_core (implicit_surface.py:218-245): surface point, its normal (third MLP
        pass) and surface_patch_warp2 (projector.py:560-645) in one C call.  The 3x3 camera matrices are formed on the
        host with the reference's own torch ops."""
        lib = _lib.load()
        dev = rays_o.device
        B = rays_o.shape[0]
        V = scene.n_src_views
        # the cameras are the ones installed in the scene (its warp maps belong to them): host copies exist already and
        # the 3x3 algebra is cached per installed view set — no device sync, no per-call inverse
        cached = getattr(scene, "_extras_params", None)
        if cached is not None and cached[0] is scene._view_key:
            p = cached[1]
        else:
            intr = scene.intrs_host if scene.intrs_host is not None else intrs.detach().float().cpu()
            c2w = scene.c2ws_host if scene.c2ws_host is not None else c2ws.detach().float().cpu()
            p = _lib.ExtrasParams()
            R0t = c2w[0, :3, :3].permute(1, 0).contiguous()
            t0 = -torch.matmul(R0t, c2w[0, :3, 3, None])
            K_inv = torch.inverse(intr)
            R_src = c2w[1:, :3, :3].permute(0, 2, 1).contiguous()
            R_rel = torch.matmul(R_src, c2w[0, :3, :3])
            RC = torch.matmul(R_src, (c2w[0, :3, 3][None, ...] - c2w[1:, :3, 3])[..., None])
            p.R0t[:] = R0t.reshape(-1).tolist()
            p.t0[:] = t0.reshape(-1).tolist()
            p.K0[:] = intr[0, :3, :3].reshape(-1).tolist()
            p.K0inv[:] = K_inv[0, :3, :3].reshape(-1).tolist()
            for v in range(V):
                p.Ksrc[v][:] = intr[v + 1, :3, :3].reshape(-1).tolist()
                p.Rrel[v][:] = R_rel[v].reshape(-1).tolist()
                p.RC[v][:] = RC[v].reshape(-1).tolist()
            p.n_src = V
            p.patch_size = PATCH_SIZE
            scene._extras_params = (scene._view_key, p)
        npx = PATCH_SIZE * PATCH_SIZE
        f32 = dict(dtype=torch.float32, device=dev)
        pts0 = torch.empty((B, 3), **f32)
        nrm0 = torch.empty((B, 3), **f32)
        ref_val = torch.empty((1, B, npx, 12), **f32)
        src_val = torch.empty((V, B, npx, 12), **f32)
        with torch.cuda.device(dev):
            ws_bytes = int(lib.surf_extras_workspace_bytes(B))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.surf_render_extras(scene.handle, self.net_handle(), C.byref(p), rays_o.data_ptr(),
                                              rays_d.data_ptr(), t["z_cross"].data_ptr(), t["z_max"].data_ptr(), B,
                                              pts0.data_ptr(), nrm0.data_ptr(), ref_val.data_ptr(), src_val.data_ptr(),
                                              ws.data_ptr(), ws_bytes, int(self.mlp_mode), _stream()), "render_extras")
        return {"ref_gray_val": ref_val, "sampled_gray_val": src_val, "_pts_sdf0": pts0, "_normal_sdf0": nrm0}

    def _extras(self, t, scene, rays_o, rays_d, intrs, c2ws):
        """ref_gray_val / sampled_gray_val of the 18-key dict (the training extras); the cameras come from the call's
        arguments, or from the prepared scene when the caller passed a PreparedScene and no matrices."""
        if intrs is None or c2ws is None:
            intrs, c2ws = scene.intrs_host, scene.c2ws_host
        if intrs is None or not scene.has_images:
            return {}
        ro = rays_o.detach().to(torch.float32).contiguous()
        rd = rays_d.detach().to(torch.float32).contiguous()
        ret = self._render_extras(t, scene, ro, rd, intrs, c2ws)
        # smooth_error (:172): |Hessian . 1| of the evaluated samples (0 elsewhere, :99), weighted by inside_sphere
        B, S = t["mid_z_vals"].shape
        pts = (ro[:, None, :] + rd[:, None, :] * t["mid_z_vals"][..., None]).reshape(-1, 3)
        smooth = self.sdf_network.smooth(pts, scene, flags=t["point_flags"])
        inside = t["inside_sphere"]
        ret["smooth_error"] = (torch.linalg.norm(smooth, ord=2, dim=-1).reshape(B, S) * inside).sum() / (inside.sum() + 1e-5)
        ret["_smooth"] = smooth
        return ret

    def _finish_dict(self, t, scene, B, S, pts_random, device):
        inv_s = self.deviation_network.inv_s().to(device)
        ge = t["gradient_error_sums"]
        ret = {
            "color_fine": t["color_fine"],
            "render_depth": t["render_depth"],
            "sdf_depth": t["sdf_depth"],
            "normal": t["normal"],
            "valid_mask": t["valid_mask"].bool(),
            "mid_z_vals": t["mid_z_vals"],
            "gradients": t["gradients"],
            "s_val": (1.0 / inv_s).reshape(1, 1).expand(B * S, 1),
            "weights": t["weights"],
            "weight_sum": t["weight_sum"],
            "weight_max": t["weight_max"],
            "gradient_error": ge[0] / (ge[1] + 1e-5),
            "inside_sphere": t["inside_sphere"],
            "mid_inside_sphere": t["mid_inside_sphere"],
        }
        if pts_random is not None:
            ret["sparse_sdf"] = torch.cat([self._sparse_sdf_random(scene, pts_random, device), t["sdf"]])
        else:
            ret["sparse_sdf"] = t["sdf"]
        return ret

    # -- reference API -----------------------------------------------------------------------------
    def render_core(self, rays_o, rays_d, z_vals, sample_dist, volumes, sparse_idxes, mask_volumes, features,
                    match_features, imgs, intrs, c2ws, near, far, cos_anneal_ratio, step, pts_random=None,
                    return_stages=False, extras=True):
        """implicit_surface.py:64-266 (inference keys).  ``sample_dist`` must be 2/n_samples[0]."""
        scene = volumes if isinstance(volumes, PreparedScene) else self.prepare(
            None, volumes, sparse_idxes, mask_volumes, imgs, features, intrs, c2ws)
        B, S = z_vals.shape
        if pts_random is None:
            pts_random = torch.rand([N_RANDOM_PTS, 3]) * 2 - 1          # same draw as the reference (Q1)
        t = self._render_device(scene, rays_o, rays_d, None, None, None, z_vals, cos_anneal_ratio, 0,
                                stages=return_stages)
        ret = self._finish_dict(t, scene, B, S, pts_random, rays_o.device)
        if extras:
            ret.update(self._extras(t, scene, rays_o, rays_d, intrs, c2ws))
        if return_stages:
            for k in ("point_flags", "point_color", "point_views", "prev_idx", "alpha"):
                ret["_" + k] = t[k]
        return ret

    def render(self, rays_o, rays_d, near, far, matching_volume, volumes, sparse_idxes, mask_volumes, imgs, features,
               match_features, intrs, c2ws, cos_anneal_ratio, step, t_rand=None, pts_random=None, return_stages=False,
               extras=True):
        """implicit_surface.py:268-335.  With ``t_rand`` / ``pts_random`` = None the random numbers are drawn
        from torch's global CPU generator in the reference's order, so seeding reproduces its sample positions."""
        scene = matching_volume if isinstance(matching_volume, PreparedScene) else self.prepare(
            matching_volume, volumes, sparse_idxes, mask_volumes, imgs, features, intrs, c2ws)
        B, S = rays_o.shape[0], sum(self.n_samples)
        if near.shape[0] == 1 and B != 1:
            near, far = near.expand(B, 1), far.expand(B, 1)
        if self.perturb > 0 and t_rand is None:
            t_rand = self.draw_t_rand(B, len(self.n_samples))
        if pts_random is None:
            pts_random = torch.rand([N_RANDOM_PTS, 3]) * 2 - 1
        t = self._render_device(scene, rays_o, rays_d, near, far, t_rand, None, cos_anneal_ratio, 0,
                                stages=return_stages)
        ret = self._finish_dict(t, scene, B, S, pts_random, rays_o.device)
        if extras:
            ret.update(self._extras(t, scene, rays_o, rays_d, intrs, c2ws))
        if return_stages:
            for k in ("point_flags", "point_color", "point_views", "prev_idx", "alpha"):
                ret["_" + k] = t[k]
        return ret

    def sdf_grid(self, scene, bound_min, bound_max, resolution, x_range=None, sparsify=False, fill=-100.0):
        """u = -sdf on the extract_geometry grid (implicit_surface.py:339-351), optionally only the x-slab
        ``x_range=(x0,x1)``.  Coordinates come from host ``torch.linspace`` exactly like the reference."""
        lib = _lib.load()
        dev = scene.device
        bmin = [float(v) for v in bound_min]
        bmax = [float(v) for v in bound_max]
        X = torch.linspace(bmin[0], bmax[0], resolution)
        Y = torch.linspace(bmin[1], bmax[1], resolution)
        Z = torch.linspace(bmin[2], bmax[2], resolution)
        if x_range is not None:
            X = X[x_range[0]:x_range[1]]
        X, Y, Z = X.to(dev), Y.to(dev), Z.to(dev)
        u = torch.empty((X.numel(), resolution, resolution), dtype=torch.float32, device=dev)
        net = self.net_handle()
        # keep every launch below 2^31 points
        max_planes = max(1, (2 ** 31 - 1) // (resolution * resolution))
        with torch.cuda.device(dev):
            for x0 in range(0, X.numel(), max_planes):
                xs = X[x0:x0 + max_planes].contiguous()
                _lib.check(lib.surf_sdf_grid(scene.handle, net, xs.data_ptr(), xs.numel(), Y.data_ptr(), resolution,
                                             Z.data_ptr(), resolution, u[x0:x0 + xs.numel()].data_ptr(),
                                             1 if sparsify else 0, float(fill), int(self.mlp_mode), _stream()),
                           "sdf_grid")
        return u

    def extract_geometry(self, volumes, sparse_idxes, bound_min, bound_max, resolution, threshold):
        """implicit_surface.py:337-357.  The SDF grid is evaluated on the GPU in one pass and meshed on the GPU
        (surf_b200/csrc/marching.cu) — the grid never visits the host; the reference's PyMCubes host call is an
        un-vendored third party.  Returns numpy (vertices float64 in world units, triangles int64)."""
        scene = volumes if isinstance(volumes, PreparedScene) else self.scene_cache.get(volumes, sparse_idxes)
        u = self.sdf_grid(scene, bound_min, bound_max, resolution)
        v, t = mesh.marching_cubes_device(u, float(threshold))
        vertices, triangles = v.cpu().numpy(), t.cpu().numpy().astype(np.int64)
        b_max_np = np.asarray([float(v_) for v_ in bound_max], dtype=np.float32)
        b_min_np = np.asarray([float(v_) for v_ in bound_min], dtype=np.float32)
        vertices = vertices / (resolution - 1.0) * (b_max_np - b_min_np)[None, :] + b_min_np[None, :]
        return vertices, triangles

    def validate(self, rays_o, rays_d, near, far, matching_volume, volumes, sparse_idxes, mask_volumes, imgs, features,
                 match_features, intrs, c2ws, bound_min, bound_max, hw, cos_anneal_ratio=1.0, step=None,
                 extract_geometry=True, mesh_resolution=512, threshold=0.0, device_outputs=False):
        """implicit_surface.py:359-402.  Same outputs (CPU tensor / numpy arrays); the 256-ray chunk loop of
        the reference becomes a few large launches that keep its per-chunk semantics (RNG stream, Q1; empty-mask
        fallback, Q6)."""
        scene = matching_volume if isinstance(matching_volume, PreparedScene) else self.prepare(
            matching_volume, volumes, sparse_idxes, mask_volumes, imgs, features, intrs, c2ws)
        outputs = {}
        if extract_geometry:
            vertices, triangles = self.extract_geometry(scene, None, bound_min, bound_max, mesh_resolution, threshold)
            outputs["vertices"] = vertices
            outputs["triangles"] = triangles
        height, width = int(hw[0]), int(hw[1])
        res = self.render_image(scene, rays_o, rays_d, near, far, cos_anneal_ratio)
        color_fine = res["color_fine"]
        normals = res["val_normal"]
        if device_outputs:
            outputs.update(res)
            return outputs
        color_fine = color_fine.cpu()
        img_fine = (color_fine.numpy().reshape([height, width, 3]) * 256).clip(0, 255)
        normal_img = normals.cpu().numpy()
        rot = np.linalg.inv(c2ws[0, :3, :3].detach().cpu().numpy())
        normal_img = (np.matmul(rot[None, :, :], normal_img[:, :, None]).reshape([height, width, 3]) * 128 + 128).clip(0, 255)
        outputs["color_fine"] = color_fine
        outputs["img_fine"] = img_fine
        outputs["normal_img"] = normal_img
        outputs["sdf_depth"] = res["sdf_depth"].cpu().numpy().reshape([height, width])
        outputs["render_depth"] = res["render_depth"].cpu().numpy().reshape([height, width])
        return outputs

    def draw_chunk_randoms(self, n_rays, chunk=256):
        """All jitter draws of a chunked validation pass in the reference's stream order (Q1): per 256-ray
        chunk 4 x rand([b,1]) then rand([1024,3]) (render_core's random points, consumed even in val).
        Returns t_rand (n_rays, 4)."""
        n_st = len(self.n_samples)
        full, rem = divmod(n_rays, chunk)
        per = n_st * chunk + N_RANDOM_PTS * 3
        parts = []
        if full:
            r = torch.rand(full * per).reshape(full, per)[:, :n_st * chunk]
            parts.append(r.reshape(full, n_st, chunk).permute(0, 2, 1).reshape(full * chunk, n_st))
        if rem:
            r = torch.rand(n_st * rem + N_RANDOM_PTS * 3)[:n_st * rem]
            parts.append(r.reshape(n_st, rem).t())
        return torch.cat(parts, dim=0).contiguous()

    def render_image(self, scene, rays_o, rays_d, near, far, cos_anneal_ratio=1.0, chunk=256, t_rand=None):
        """The image part of validate(): device tensors color_fine (n,3), val_normal (n,3),
        sdf_depth (n,1), render_depth (n,)."""
        n = rays_o.shape[0]
        if near.shape[0] == 1 and n != 1:
            near, far = near.expand(n, 1), far.expand(n, 1)
        draw = self.perturb > 0 and t_rand is None
        if t_rand is not None:
            t_rand = t_rand.to(rays_o.device, non_blocking=True)
        step_rays = max(chunk, (self.ray_batch // chunk) * chunk)
        acc = {k: [] for k in ("color_fine", "val_normal", "sdf_depth", "render_depth")}
        if n == 0:      # an empty shard (more ranks than 256-ray chunks): correctly shaped empty tensors
            f32 = dict(dtype=torch.float32, device=rays_o.device)
            return {"color_fine": torch.empty((0, 3), **f32), "val_normal": torch.empty((0, 3), **f32),
                    "sdf_depth": torch.empty((0, 1), **f32), "render_depth": torch.empty((0,), **f32)}
        for r0 in range(0, n, step_rays):
            r1 = min(n, r0 + step_rays)
            tr = None if t_rand is None else t_rand[r0:r1]
            if draw:
                # the reference's jitter stream is drawn on the host, batch by batch: launches are asynchronous, so
                # the draw for this batch overlaps the kernels of the previous one (same stream order: chunks ascend)
                tr = self.draw_chunk_randoms(r1 - r0, chunk).pin_memory().to(rays_o.device, non_blocking=True)
            t = self._render_device(scene, rays_o[r0:r1], rays_d[r0:r1], near[r0:r1], far[r0:r1],
                                    tr, None, cos_anneal_ratio, chunk, lean=True)
            for k in acc:
                acc[k].append(t[k])
        return {k: (v[0] if len(v) == 1 else torch.cat(v, dim=0)) for k, v in acc.items()}

    def forward(self, mode, ipts, matching_volume, volumes, sparse_idxes, mask_volumes, features, match_features,
                cos_anneal_ratio=1.0, step=None):
        """implicit_surface.py:404-436."""
        imgs, intrs, c2ws = ipts["imgs"], ipts["intrs"], ipts["c2ws"]
        rays_o, rays_d, near, far = ipts["rays_o"], ipts["rays_d"], ipts["near"], ipts["far"]
        if near.shape[0] == 1:
            near = near.repeat(rays_o.shape[0], 1)
            far = far.repeat(rays_o.shape[0], 1)
        if isinstance(matching_volume, PreparedScene):      # a scene built in the compact layout (Volume.to_prepared_scene)
            scene = matching_volume.set_views(imgs, features, intrs, c2ws)
        else:
            scene = self.prepare(matching_volume, volumes, sparse_idxes, mask_volumes, imgs, features, intrs, c2ws)
        if mode == "val":
            # the reference hard-wires extract_geometry=True, mesh_resolution=512, threshold=0 here (:416-418);
            # ``val_options`` (set by surf_b200.runner.validate) may override them, e.g. a smaller mesh for previews
            outputs = self.validate(rays_o, rays_d, near, far, scene, None, None, None, imgs, features, match_features,
                                    intrs, c2ws, ipts["bound_min"], ipts["bound_max"], ipts["hw"], cos_anneal_ratio, step,
                                    **getattr(self, "val_options", {}))
        else:
            outputs = self.render(rays_o, rays_d, near, far, scene, None, None, None, imgs, features, match_features,
                                  intrs, c2ws, cos_anneal_ratio, step)
        if "pseudo_pts" in ipts:
            p = ipts["pseudo_pts"].to(rays_o.device)
            outputs["pseudo_sdf"] = self._sparse_sdf_random(scene, p, rays_o.device)
        return outputs
