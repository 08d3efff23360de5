"""Generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference or $SURF_REF):

    python oracle/make_golden.py

The reference has no golden vectors of its own, so these are
outputs of the reference itself (imported from where it lies through
oracle/ref_loader.py — nothing is copied) on small synthetic scenes built by
surf_b200/synthetic.py.  Each file stores the network state_dict, the ray
inputs, the scene recipe (regenerated from seeds at test time; a checksum of
the scene tensors guards against generator drift) and every output tensor.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from surf_b200 import synthetic  # noqa: E402
from surf_b200.conf import default_implicit_surface_conf  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def scene_checksum(sc):
    h = hashlib.sha256()
    for t in [sc.imgs, sc.intrs, sc.c2ws, sc.near, sc.far, sc.matching_volume] + sc.volumes + sc.sparse_idxes \
            + sc.mask_volumes + sc.features:
        h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


def build_reference_net(IS, weight_seed, perturb_weights, variance=None, scene=None):
    conf = default_implicit_surface_conf()
    torch.manual_seed(weight_seed)
    net = IS.ImplicitSurface(conf)
    if perturb_weights:
        # geometric init zeroes every feature / PE-frequency column (sdf_network.py:71-86), which
        # would leave the sparse-volume path untested: add noise to all parameters.
        g = torch.Generator().manual_seed(weight_seed + 1)
        with torch.no_grad():
            for name, p in net.named_parameters():
                if name.endswith("weight_v"):
                    zero_cols = (p == 0).to(p.dtype)       # feature / PE-frequency columns
                    p.add_(torch.randn(p.shape, generator=g) * (0.01 + 0.04 * zero_cols))
                elif name.endswith("weight_g"):
                    p.mul_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                elif name.endswith("bias"):
                    p.add_(torch.randn(p.shape, generator=g) * 0.02)
        if scene is not None:
            # the noise shifts the level set; re-centre it on the r=0.5 sphere so rays still cross a surface
            q = torch.randn(512, 3, generator=g)
            q = 0.5 * q / q.norm(dim=-1, keepdim=True)
            with torch.no_grad():
                shift = net.sdf_network.sdf(q, scene.volumes, scene.sparse_idxes).mean()
                net.sdf_network.lin6.bias[0] -= shift
    if variance is not None:
        with torch.no_grad():
            net.deviation_network.variance.fill_(variance)
    net.eval()
    return net


def to_np(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return np.asarray(v)


def save(name, recipe, net, inputs, outputs):
    d = {}
    for k, v in recipe.items():
        d["recipe." + k] = np.asarray(v)
    for k, v in net.state_dict().items():
        d["sd." + k] = to_np(v)
    for k, v in inputs.items():
        d["in." + k] = to_np(v)
    for k, v in outputs.items():
        d["out." + k] = to_np(v)
    d["torch_version"] = np.asarray(torch.__version__)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def case_render(IS, name, nv, H, W, base, scene_seed, n_rays, ray_seed, weight_seed, perturb_weights,
                variance, torch_seed, miss=False):
    sc = synthetic.make_scene(nv, H, W, base, seed=scene_seed)
    net = build_reference_net(IS, weight_seed, perturb_weights, variance, sc)
    o, d = synthetic.random_pixel_rays(sc, n_rays, seed=ray_seed)
    if miss:
        # rays that leave the volume immediately: exercises the empty-mask fallback (Q6)
        o = o + torch.tensor([0.0, 0.0, -6.0])
    near = sc.near.expand(n_rays, 1).contiguous()
    far = sc.far.expand(n_rays, 1).contiguous()
    torch.manual_seed(torch_seed)
    out = net.render(o, d, near, far, *sc.render_args(), 1.0, None)
    # stage outputs recomputed from the reference's own sub-functions on the same inputs
    P = IS  # module namespace
    mid_z = out["mid_z_vals"]
    pts = (o[:, None, :] + d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    vmask = P.lookup_volume(pts, sc.mask_volumes, sample_mode="nearest").any(dim=-1)
    extra = {"_voxel_mask": vmask}
    sel = vmask.clone()
    if int(sel.sum()) < 1:
        sel[:10] = True
    pv = pts[sel]
    from models.modules.projector import lookup_sparse_volume
    extra["_pts_valid"] = pv
    extra["_sparse_feats"] = lookup_sparse_volume(pv.clone(), sc.volumes, sc.sparse_idxes)
    fv, rd, mv = P.lookup_feature(pv, sc.imgs, sc.intrs, sc.c2ws, sc.features)
    extra["_feat_views"], extra["_ray_diff"], extra["_view_mask"] = fv, rd, mv
    extra["_blend_rgb"] = net.color_network(fv.clone(), rd, mv)
    extra["_sdf_full"] = net.sdf_network(pv, sc.volumes, sc.sparse_idxes)
    gr, sm = net.sdf_network.gradient(pv.clone(), sc.volumes, sc.sparse_idxes)
    extra["_grad_valid"], extra["_smooth_valid"] = gr, sm
    out = {k: v for k, v in out.items()}
    out.update(extra)
    recipe = dict(nv=nv, H=H, W=W, base=base, scene_seed=scene_seed, torch_seed=torch_seed,
                  scene_sha=scene_checksum(sc))
    save(name, recipe, net, {"rays_o": o, "rays_d": d, "near": near, "far": far}, out)


def case_validate(IS, name, nv, H, W, base, scene_seed, res_level, weight_seed, torch_seed):
    sc = synthetic.make_scene(nv, H, W, base, seed=scene_seed)
    net = build_reference_net(IS, weight_seed, True, None, sc)
    o, d, hw = synthetic.image_rays(sc, res_level)
    near = sc.near.expand(o.shape[0], 1).contiguous()
    far = sc.far.expand(o.shape[0], 1).contiguous()
    torch.manual_seed(torch_seed)
    out = net.validate(o, d, near, far, *sc.render_args(), torch.tensor([-1.0, -1, -1]), torch.tensor([1.0, 1, 1]),
                       hw, 1.0, None, extract_geometry=False)
    recipe = dict(nv=nv, H=H, W=W, base=base, scene_seed=scene_seed, torch_seed=torch_seed,
                  res_level=res_level, scene_sha=scene_checksum(sc))
    save(name, recipe, net, {"rays_o": o, "rays_d": d, "near": near, "far": far}, out)


def case_sdf_grid(IS, name, base, scene_seed, weight_seed, resolution):
    sc = synthetic.make_scene(3, 48, 64, base, seed=scene_seed)
    net = build_reference_net(IS, weight_seed, True, None, sc)
    # extract_geometry's loop body (implicit_surface.py:339-351) without the marching-cubes call
    bmin, bmax = torch.tensor([-1.0, -1, -1]), torch.tensor([1.0, 1, 1])
    N = 64
    X = torch.linspace(bmin[0], bmax[0], resolution).split(N)
    Y = torch.linspace(bmin[1], bmax[1], resolution).split(N)
    Z = torch.linspace(bmin[2], bmax[2], resolution).split(N)
    u = np.zeros([resolution] * 3, dtype=np.float32)
    with torch.no_grad():
        for xi, xs in enumerate(X):
            for yi, ys in enumerate(Y):
                for zi, zs in enumerate(Z):
                    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing="ij")
                    pts = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
                    val = -net.sdf_network.sdf(pts, sc.volumes, sc.sparse_idxes).reshape(len(xs), len(ys), len(zs))
                    u[xi * N: xi * N + len(xs), yi * N: yi * N + len(ys), zi * N: zi * N + len(zs)] = val.numpy()
    # points outside [-1,1]^3 (extrapolating corner weights, Q13) and exactly on voxel centres
    g = torch.Generator().manual_seed(5)
    wild = (torch.rand(256, 3, generator=g) * 3.0 - 1.5)
    n_fine = sc.sparse_idxes[0].shape[0]
    centres = torch.randint(0, n_fine, (64, 3), generator=g).float() * (2.0 / (n_fine - 1)) - 1.0
    wild = torch.cat([wild, centres, torch.tensor([[-1.0, -1, -1], [1, 1, 1], [0, 0, 0]])])
    with torch.no_grad():
        full = net.sdf_network(wild, sc.volumes, sc.sparse_idxes)
    gr, sm = net.sdf_network.gradient(wild.clone(), sc.volumes, sc.sparse_idxes)
    recipe = dict(nv=3, H=48, W=64, base=base, scene_seed=scene_seed, resolution=resolution,
                  scene_sha=scene_checksum(sc))
    save(name, recipe, net, {"wild_pts": wild}, {"u": u, "wild_full": full, "wild_grad": gr, "wild_smooth": sm})


def case_matching_field(IS, name, nv, H, W, base, scene_seed, torch_seed):
    """MatchingField.forward (matching_field.py:74-141) of the unmodified reference: stage 0 (full range), stage 1
    (two windows around the stage-0 depths), and a jittered stage-1 pass (RNG stream: rand([B,1]) per window)."""
    from models.modules.matching_field import MatchingField
    sc = synthetic.make_scene(nv, H, W, base, seed=scene_seed)
    conf = ref_loader.DictConf({"n_samples_depths": [128, 64, 32, 16], "n_importance_depths": [128, 64, 32, 16],
                                "up_sample_steps": [4, 4, 4, 4], "depth_res_levels": [4, 2, 2, 1]})
    mf = MatchingField(conf)
    near_fars = torch.stack([torch.tensor([float(sc.near), float(sc.far)])] * nv)
    ipts = {"near_fars": near_fars, "c2ws": sc.c2ws, "intrs": sc.intrs, "imgs": sc.imgs, "src_idx": 1}
    rr = [1.0, 0.4, 0.1, 0.01]
    out = {}
    with torch.no_grad():
        d0, o0 = mf(ipts, sc.matching_volume, 0, rr, None, perturb=False)
        d1, o1 = mf(ipts, sc.matching_volume, 1, rr, d0, perturb=False)
        torch.manual_seed(torch_seed)
        d1p, o1p = mf(ipts, sc.matching_volume, 1, rr, d0, perturb=True)
        d3, o3 = mf(ipts, sc.matching_volume, 3, rr, d1, perturb=False)
    for tag, (dd, oo) in {"s0": (d0, o0), "s1": (d1, o1), "s1p": (d1p, o1p), "s3": (d3, o3)}.items():
        out["depth_" + tag] = torch.stack(dd)
        out["occ_" + tag] = torch.stack(oo)
    recipe = dict(nv=nv, H=H, W=W, base=base, scene_seed=scene_seed, torch_seed=torch_seed, scene_sha=scene_checksum(sc))
    # no network in this case: save an empty state dict through the common writer
    class _NoNet:
        def state_dict(self):
            return {}
    save(name, recipe, _NoNet(), {"near_fars": near_fars}, out)


def case_volume(IS, name, nv, H, W, base, scene_seed, weight_seed):
    """The producers of the scene tensors (volume.py:21-168) of the unmodified reference on synthetic inputs: two
    stages of init_coords / back_proj_multiscale / sparse2dense / get_index / up_sample / depth_filtering."""
    from models.modules.volume import Volume
    sc = synthetic.make_scene(nv, H, W, base, seed=scene_seed)
    torch.manual_seed(weight_seed)
    vol = Volume(ref_loader.DictConf({"base_volume_dim": [base, base, base]}))
    feats = [f for f in sc.features[::-1]]            # coarse -> fine, as feature_network emits them
    g = torch.Generator().manual_seed(weight_seed + 1)
    out = {}
    with torch.no_grad():
        c0 = vol.init_coords().type_as(sc.intrs)
        fv0, m0 = vol.back_proj_multiscale(feats, c0, sc.intrs, sc.c2ws, 0)
        c0m, fv0m = c0[m0], fv0[m0]
        reg0 = torch.randn(c0m.shape[0], 8, generator=g)                      # stand-in for reg_network's output
        mv0, mk0 = vol.sparse2dense(reg0[:, :1], c0m, None)
        idx0 = vol.get_index(c0m)
        depths = [torch.full((H, W), 2.0) + 0.05 * torch.randn(H, W, generator=g) for _ in range(nv)]
        c1, f1 = vol.up_sample(c0m.clone(), reg0)
        keep = None
        c1f, f1f = vol.depth_filtering(depths, c1, f1, sc.intrs, sc.c2ws, 0.4)
        fv1, m1 = vol.back_proj_multiscale(feats, c1f, sc.intrs, sc.c2ws, 1)
        c1m = c1f[m1]
        reg1 = torch.randn(c1m.shape[0], 8, generator=g)
        mv1, mk1 = vol.sparse2dense(reg1[:, :1], c1m, mv0)
        idx1 = vol.get_index(c1m)
    out.update({"c0": c0, "fv0": fv0, "m0": m0, "reg0": reg0, "mv0": mv0, "mk0": mk0, "idx0": idx0,
                "depths": torch.stack(depths), "c1": c1, "f1": f1, "c1f": c1f, "f1f": f1f, "fv1": fv1, "m1": m1,
                "reg1": reg1, "mv1": mv1, "mk1": mk1, "idx1": idx1})
    recipe = dict(nv=nv, H=H, W=W, base=base, scene_seed=scene_seed, weight_seed=weight_seed, scene_sha=scene_checksum(sc))

    class _Agg:
        def state_dict(self_inner):
            return {"agg_mlp." + k: v for k, v in vol.agg_mlp.state_dict().items()}
    save(name, recipe, _Agg(), {}, out)


def case_fpn(IS, name, nv, H, W, weight_seed, img_seed):
    """FeatureNetwork.forward (feature_network.py:126-178) of the unmodified reference: 4-stage encoder / decoder with
    InstanceNorm, d_base 8, d_out [4,4,4,4] (confs/surf.conf)."""
    from models.modules.feature_network import FeatureNetwork
    conf = ref_loader.DictConf({"d_in": 3, "d_base": 8, "d_out": [4, 4, 4, 4]})
    torch.manual_seed(weight_seed)
    net = FeatureNetwork(conf).eval()
    imgs = torch.rand((nv, 3, H, W), generator=torch.Generator().manual_seed(img_seed))
    with torch.no_grad():
        outs = net(imgs)
    save(name, dict(nv=nv, H=H, W=W, weight_seed=weight_seed, img_seed=img_seed), net, {"imgs": imgs},
         {"feat%d" % i: o for i, o in enumerate(outs)})


def load_reference_surf():
    """models/surf.py of the unmodified reference; `torchsparse` (2.1.0, un-vendored) is shimmed with a container
    SparseTensor and inert layer classes — the regularisation network itself is replaced by oracle/standin_reg.py."""
    import types
    import torch.nn as nn
    if "torchsparse" not in sys.modules:
        ts, tsn, tst = types.ModuleType("torchsparse"), types.ModuleType("torchsparse.nn"), types.ModuleType("torchsparse.tensor")

        class SparseTensor:
            def __init__(self, feats, coords):
                self.F, self.C = feats, coords
        tst.SparseTensor = SparseTensor
        for name in ("Conv3d", "BatchNorm", "ReLU"):
            setattr(tsn, name, lambda *a, **k: nn.Identity())
        ts.nn, ts.tensor = tsn, tst
        sys.modules.update({"torchsparse": ts, "torchsparse.nn": tsn, "torchsparse.tensor": tst})
    from models import surf as ref_surf
    return ref_surf


def case_build_volumes(IS, name, nv, H, W, base, scene_seed, weight_seed):
    """SuRF.build_volumes (surf.py:80-131) of the unmodified reference — FeatureNetwork, Volume, MatchingField and the
    four-stage orchestration — with the stand-in regulariser of oracle/standin_reg.py in place of the torchsparse net."""
    import standin_reg
    ref_surf = load_reference_surf()
    sc = synthetic.make_scene(nv, H, W, base, seed=scene_seed)
    model_conf = ref_loader.DictConf({
        "range_ratios": [1.0, 0.4, 0.1, 0.01],
        "feature_network": ref_loader.DictConf({"d_in": 3, "d_base": 8, "d_out": [4, 4, 4, 4]}),
        "volume": ref_loader.DictConf({"base_volume_dim": [base, base, base]}),
        "reg_network": ref_loader.DictConf({"d_in": [8, 16, 16, 16], "d_base": [8, 8, 8, 8], "d_out": [8, 8, 8, 8]}),
        "matching_field": ref_loader.DictConf({"n_samples_depths": [128, 64, 32, 16], "n_importance_depths": [128, 64, 32, 16],
                                               "up_sample_steps": [4, 4, 4, 4], "depth_res_levels": [4, 2, 2, 1]}),
        "implicit_surface": default_implicit_surface_conf()})
    torch.manual_seed(weight_seed)
    model = ref_surf.SuRF(model_conf).eval()
    model.reg_network = standin_reg.StandinReg()
    near_fars = torch.stack([torch.tensor([float(sc.near), float(sc.far)])] * nv)
    ipts = {"imgs": sc.imgs, "intrs": sc.intrs, "c2ws": sc.c2ws, "near": sc.near, "far": sc.far, "near_fars": near_fars,
            "src_idx": 1}
    with torch.no_grad():
        features = model.feature_network(sc.imgs)
        outputs, volumes, idxs, masks, matching = model.build_volumes(ipts, features, False)
    out = {"matching_volume": matching}
    for s in range(4):
        out["volume%d" % s] = volumes[s]
        out["sparse_idx%d" % s] = idxs[s].to(torch.int32)
        out["mask%d" % s] = masks[s].to(torch.uint8)
        out["depth_stage%d" % s] = outputs["depth_stage%d" % s]
        out["depth_src_stage%d" % s] = outputs["depth_src_stage%d" % s]
    recipe = dict(nv=nv, H=H, W=W, base=base, scene_seed=scene_seed, weight_seed=weight_seed, scene_sha=scene_checksum(sc))

    class _Nets:
        def state_dict(self_inner):
            d = {"feature_network." + k: v for k, v in model.feature_network.state_dict().items()}
            d.update({"volume." + k: v for k, v in model.volume.state_dict().items()})
            return d
    save(name, recipe, _Nets(), {"features%d" % i: f for i, f in enumerate(features)}, out)


def main():
    os.makedirs(OUT, exist_ok=True)
    IS = ref_loader.load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    # A: DTU-shaped val (3 views), perturbed weights so every path matters
    case_render(IS, "render_v2_perturbed", nv=3, H=48, W=64, base=8, scene_seed=1, n_rays=48, ray_seed=2,
                weight_seed=0, perturb_weights=True, variance=None, torch_seed=0)
    # B: pristine geometric init (sphere SDF), larger inv_s
    case_render(IS, "render_v2_init", nv=3, H=48, W=64, base=8, scene_seed=3, n_rays=32, ray_seed=4,
                weight_seed=0, perturb_weights=False, variance=0.8, torch_seed=7)
    # C: training-shaped (5 views)
    case_render(IS, "render_v4_perturbed", nv=5, H=60, W=80, base=8, scene_seed=11, n_rays=32, ray_seed=12,
                weight_seed=3, perturb_weights=True, variance=0.5, torch_seed=5)
    # D: rays missing the volume -> empty-mask fallback
    case_render(IS, "render_miss", nv=3, H=48, W=64, base=8, scene_seed=1, n_rays=16, ray_seed=6,
                weight_seed=0, perturb_weights=True, variance=None, torch_seed=1, miss=True)
    # E: chunked validation image (3 chunks of 256 rays -> RNG stream across chunks, Q1)
    case_validate(IS, "validate_24x32", nv=3, H=48, W=64, base=8, scene_seed=1, res_level=2, weight_seed=0,
                  torch_seed=0)
    # F: SDF grid + out-of-range points
    case_sdf_grid(IS, "sdf_grid_24", base=8, scene_seed=1, weight_seed=0, resolution=24)
    # G: the upstream user of the probe kernel: MatchingField.forward
    case_matching_field(IS, "matching_field", nv=3, H=48, W=64, base=8, scene_seed=1, torch_seed=9)
    # H: the producers of the scene tensors: Volume.*
    case_volume(IS, "volume", nv=3, H=48, W=64, base=8, scene_seed=1, weight_seed=4)
    # I: the 2-D feature pyramid (first half)
    case_fpn(IS, "fpn", nv=2, H=40, W=56, weight_seed=5, img_seed=6)
    # J: the four-stage volume construction around a stand-in regulariser (A13: SuRF.init_volumes / build_volumes)
    case_build_volumes(IS, "build_volumes", nv=3, H=48, W=64, base=8, scene_seed=1, weight_seed=8)


if __name__ == "__main__":
    main()
