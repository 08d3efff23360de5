"""Shared helpers for the parity tests: golden loading, scene rebuild, comparators."""
import hashlib
import os

import numpy as np
import torch

from surf_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp32 tolerance of BASELINE.json's north_star: 1e-4 relative.  "Relative" is taken against the
# scale of the compared tensor (SDF crosses zero, so a pure element-wise ratio is meaningless):
#   |a-b| <= RTOL * max(|b|_elem, scale(b))   with scale = max|b|
RTOL_FP32 = 1e-4
RTOL_BF16 = 1e-2


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    g = {"recipe": {}, "sd": {}, "in": {}, "out": {}}
    for k in z.files:
        if "." in k and k.split(".", 1)[0] in g:
            grp, rest = k.split(".", 1)
            g[grp][rest] = z[k]
    g["sd"] = {k: torch.from_numpy(np.array(v)) for k, v in g["sd"].items()}
    g["in"] = {k: torch.from_numpy(np.array(v)) for k, v in g["in"].items()}
    return g


def scene_checksum(sc):
    h = hashlib.sha256()
    for t in [sc.imgs, sc.intrs, sc.c2ws, sc.near, sc.far, sc.matching_volume] + sc.volumes + sc.sparse_idxes \
            + sc.mask_volumes + sc.features:
        h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


def scene_from_recipe(r, check=True):
    sc = synthetic.make_scene(int(r["nv"]), int(r["H"]), int(r["W"]), int(r["base"]), seed=int(r["scene_seed"]))
    if check:
        assert scene_checksum(sc) == str(r["scene_sha"]), "synthetic scene generator drifted from the golden recipe"
    return sc


def rel_err(a, b):
    """max |a-b| / max|b|  (scale-relative error)"""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    if b.numel() == 0:
        return 0.0
    scale = max(float(b.abs().max()), 1e-12)
    return float((a - b).abs().max()) / scale


def assert_close(a, b, rtol, what="", floor=None):
    a = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).detach().cpu().double()
    b = torch.as_tensor(np.asarray(b) if not isinstance(b, torch.Tensor) else b).detach().cpu().double()
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    if b.numel() == 0:
        return
    scale = float(b.abs().max()) if floor is None else max(float(b.abs().max()), floor)
    tol = rtol * torch.clamp(b.abs(), min=max(scale, 1e-30))
    bad = (a - b).abs() > tol
    assert not bool(bad.any()), "%s: %d/%d beyond rtol=%g (max abs err %.3e, scale %.3e)" % (
        what, int(bad.sum()), b.numel(), rtol, float((a - b).abs().max()), scale)


def assert_equal_int(a, b, what=""):
    a = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).detach().cpu()
    b = torch.as_tensor(np.asarray(b) if not isinstance(b, torch.Tensor) else b).detach().cpu()
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    ne = (a.to(torch.int64) != b.to(torch.int64))
    assert not bool(ne.any()), "%s: %d/%d integer mismatches" % (what, int(ne.sum()), b.numel())


def blend_envelope(O, net, fv, rd, mk, n_random=4, seed=0, ulps=(1, 3)):
    """Conditioning envelope of the reference's anti-alias pooling weights.

    weight = (exp(s (dot_v - 1)) - min_v exp(...)) / (sum + 1e-8) subtracts nearly equal fp32 exponentials
    (blending_network.py:76-80).  With 3+ source views of similar viewing angle a ONE-ulp change of one
    exponential moves the output by up to 0.2 in RGB (measured on the 5-view golden case), torch's own
    CPU exp is not correctly rounded (differs from the correctly rounded value in ~2 % of arguments), the argument
    s (dot - 1) itself cancels (dot ~ 0.999), and the kernel's exp is ex2.approx(x log2 e) (<= 3 ulp), so
    no independent implementation can match the reference bit-for-bit there.  Returns (reference output,
    per-point envelope) where the envelope is the largest change of the oracle's output when the
    exponentials move by +-k ulp, k in ``ulps`` (each view alone, both signs, plus random patterns)."""
    g = torch.Generator().manual_seed(seed)
    base = O.blend(net, fv, rd, mk)
    n, V = fv.shape[0], fv.shape[1]
    env = torch.zeros(n)
    pats = []
    for k in ulps:
        for v in range(V):
            for sgn in (-k, k):
                p = torch.zeros(n, V, dtype=torch.int64)
                p[:, v] = sgn
                pats.append(p)
        for _ in range(n_random):
            pats.append((torch.randint(0, 3, (n, V), generator=g) - 1) * k)
    for p in pats:
        env = torch.maximum(env, (O.blend(net, fv, rd, mk, e_ulp=p) - base).abs().max(dim=1)[0])
    return base, env


# ------------------------------------------------------------------------------------------------
# composited outputs: 1e-4 + first-order propagation of the measured per-point deviations
# ------------------------------------------------------------------------------------------------
import surf_oracle as _O

RAY_KEYS = _O.RAY_KEYS


def composite_envelope(O, ref, sdf_g, grad_g, color_g, rays_o, rays_d, inv_s, rot0_inv, cos_anneal_ratio=1.0,
                       per_sample=False):
    """See oracle/surf_oracle.py:composite_envelope (checker logic shared with bench.py's parity block)."""
    return O.composite_envelope(ref, sdf_g, grad_g, color_g, rays_o, rays_d, inv_s, rot0_inv, cos_anneal_ratio,
                                per_sample)


def assert_within_envelope(got, ref, env, what, rtol=RTOL_FP32, k_env=1.5, floor=1e-2, rows=None):
    """|got - ref| <= rtol * max(scale(ref), floor) + k_env * env, elementwise.  ``rows``: bool mask of rays to
    compare (rays excluded by a proven boundary mismatch)."""
    a = torch.as_tensor(np.asarray(got) if not isinstance(got, torch.Tensor) else got).detach().cpu().double()
    b = torch.as_tensor(np.asarray(ref) if not isinstance(ref, torch.Tensor) else ref).detach().cpu().double()
    a, b = a.reshape(env.shape), b.reshape(env.shape)
    if rows is not None:
        a, b, env = a[rows], b[rows], env[rows]
    if b.numel() == 0:
        return 0.0
    scale = max(float(b.abs().max()), floor)
    tol = rtol * scale + k_env * env
    err = (a - b).abs()
    bad = err > tol
    assert not bool(bad.any()), "%s: %d/%d beyond %g*scale + %.1f*envelope (max err %.3e, scale %.3e, max env %.3e)" % (
        what, int(bad.sum()), b.numel(), rtol, k_env, float(err.max()), scale, float(env.max()))
    return float((err / tol).max())


def explain_mask_mismatches(O, vm_gpu, vm_ref, mid_gpu, mid_ref, rays_o, rays_d, mask_volumes, what="voxel mask"):
    """Mismatch census with proof: every sample whose voxel-mask bit differs between the two paths must
    lie closer to a rounding boundary of the nearest-voxel lookup than the two paths' sample positions differ.
    Returns the bool (B,) mask of rays WITHOUT a mismatch (the ones the float comparisons then cover)."""
    B, S = mid_ref.shape
    vm_gpu = torch.as_tensor(vm_gpu).cpu().reshape(B, S).bool()
    vm_ref = torch.as_tensor(vm_ref).cpu().reshape(B, S).bool()
    mism = vm_gpu != vm_ref
    n = int(mism.sum())
    if n:
        mid_gpu = torch.as_tensor(mid_gpu).cpu().reshape(B, S)
        pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_ref[..., None])[mism]
        dz = (mid_gpu.double() - mid_ref.double()).abs()[mism]
        dist = O.voxel_round_distance(pts, mask_volumes)
        thr = 2.0 * dz + 2.5e-7          # position difference of the two paths + ~2 ulp of a coordinate
        assert bool((dist <= thr).all()), "%s: %d of %d mismatching samples are NOT on a rounding boundary " \
            "(worst: %.3e from the boundary, allowed %.3e)" % (what, int((dist > thr).sum()), n,
                                                               float((dist - thr).max()), float(thr.max()))
        assert n <= max(2, int(2e-4 * mism.numel())), "%s: %d mismatches" % (what, n)
    return ~mism.any(dim=1)


def explain_gradient_mismatches(O, out, ref, rays_o, rays_d, sc, net, what="", rtol=RTOL_FP32, max_frac=2e-3):
    """Gradient census with proof.  The trilinear feature lookup is continuous across voxel faces, its derivative is
    not, and inside a voxel it varies at (feature difference) / voxel^2 — so the reference's OWN gradient moves with
    the last bit of the sample position (oracle.gradient_position_envelope).  Every evaluated sample is held to
        |g_gpu - g_ref| <= rtol * scale + 2 * position envelope;
    a sample beyond that must be PROVEN to sit on a voxel face of some level (closer than the two paths' measured
    sample-position difference + 2.5 ulp of the grid coordinate), where the derivative jumps; anything else fails.
    Returns (bool (B,) mask of rays without a face sample, worst error / tolerance among the compared samples)."""
    B, S = ref["mid_z_vals"].shape
    cm = ref["_compute_mask"]
    gg = out["gradients"].cpu().reshape(-1, 3).double()
    gr = ref["_grad"].double()
    scale = max(float(gr[cm].abs().max()) if bool(cm.any()) else 0.0, 1e-30)
    env = O.gradient_position_envelope(net, ref, out["mid_z_vals"], rays_o, rays_d, sc.volumes, sc.sparse_idxes)
    err = (gg - gr).abs().max(dim=1)[0]
    tol = rtol * scale + 2.0 * env
    dev = (err > rtol * scale) & cm                       # candidates: look at each one
    bad = torch.zeros_like(dev)
    if bool(dev.any()):
        dz = (out["mid_z_vals"].cpu().double() - ref["mid_z_vals"].double()).abs().reshape(-1)[dev]
        dn = rays_d.double().norm(dim=1)[:, None].expand(B, S).reshape(-1)[dev]
        margin = O.voxel_face_margin(ref["_pts"][dev], sc.sparse_idxes, 2.0 * dz * dn)
        on_face = margin < 2.5
        off = ~on_face & (err[dev] > tol[dev])
        assert not bool(off.any()), "%sgradient: %d of %d deviating samples are beyond 1e-4 + position envelope and NOT on " \
            "a voxel face (worst margin %.1f ulp, worst err %.3e vs tolerance %.3e, scale %.3e)" % (
                what, int(off.sum()), int(dev.sum()), float(margin[off].max()), float(err[dev][off].max()),
                float(tol[dev][off].min()), scale)
        bad[dev.nonzero()[:, 0][on_face]] = True         # a sample ON a face: its ray is counted, not compared
        n = int(bad.sum())
        assert n <= max(3, int(max_frac * int(cm.sum()))), "%sgradient: %d on-face samples of %d" % (what, n, int(cm.sum()))
    ok = cm & ~bad
    worst = float((err[ok] / tol[ok]).max()) if bool(ok.any()) else 0.0
    return ~bad.reshape(B, S).any(dim=1), worst


def check_composited(O, out, ref, sc, net, rays_o, rays_d, rows=None, per_sample=False, what="", grad_checked=False):
    """Per-point values at 1e-4; composited values at 1e-4 + the first-order image of the measured per-point
    deviations (helpers.composite_envelope).  ``out`` = GPU dict with stages, ``ref`` = oracle dict with stages."""
    B, S = ref["mid_z_vals"].shape
    cm = ref["_compute_mask"]
    keep = torch.ones(B, dtype=torch.bool) if rows is None else rows
    keep_p = keep[:, None].expand(B, S).reshape(-1)
    sdf_g = out["sparse_sdf"][-B * S:].cpu()
    grad_g = out["gradients"].cpu().reshape(-1, 3)
    col_g = out["_point_color"].cpu().reshape(-1, 3)
    assert_close(sdf_g[keep_p & cm], ref["_sdf"][keep_p & cm], RTOL_FP32, what + "per-point sdf")
    assert bool((sdf_g[keep_p & ~cm] == 100.0).all()), what + "masked-out samples must carry sdf = 100 (Q7)"
    if grad_checked:      # the caller ran explain_gradient_mismatches (1e-4 + position envelope, faces proven)
        pass
    else:
        assert_close(grad_g[keep_p], ref["_grad"][keep_p], RTOL_FP32, what + "per-point gradient")
    # colour: 1e-4 plus the reference's own conditioning envelope of the pooling weights
    pv = ref["_pts"][cm]
    fv, rd, mv = O.lookup_feature(pv, sc.imgs, sc.intrs, sc.c2ws, sc.features)
    _, env_p = blend_envelope(O, net, fv, rd, mv, n_random=2)
    env_pt = torch.zeros(B * S)
    env_pt[cm] = env_p
    # ... and of the projection (oracle.color_position_envelope)
    env_pos = O.color_position_envelope(net, ref, out["mid_z_vals"], rays_o, rays_d, sc.imgs, sc.intrs, sc.c2ws,
                                        sc.features).float()
    env_pt = env_pt + env_pos
    cerr = (col_g - ref["_color"].reshape(-1, 3)).abs().max(dim=1)[0]
    bad = (cerr > RTOL_FP32 + 2.0 * env_pt) & keep_p
    if bool(bad.any()):          # diagnostics of the offending points (what makes them special?)
        sel = bad[cm]
        e = torch.exp(torch.abs(net.color["s"]) * (rd[sel][..., 3] - 1)).double()
        ulp = torch.finfo(torch.float32).eps * e
        spread = (e.max(dim=1)[0] - e.min(dim=1)[0]) / ulp.max(dim=1)[0]
        border = O.projection_border_distance(pv[sel], sc.intrs, sc.c2ws, sc.features)
        msg = "; ".join("err %.2e env %.2e mask %s exp-spread %.1f ulp border %.2e px col_gpu %s col_ref %s" % (
            float(cerr[cm][sel][i]), float(env_p[sel][i]), mv[sel][i].tolist(), float(spread[i]), float(border[i]),
            [round(float(v), 5) for v in col_g[cm][sel][i]], [round(float(v), 5) for v in ref["_color"].reshape(-1, 3)[cm][sel][i]])
            for i in range(min(4, int(sel.sum()))))
        raise AssertionError(what + "per-point colour: %d points beyond 1e-4 + conditioning envelope (max %.3e): %s" % (
            int(bad.sum()), float(cerr[keep_p].max()), msg))
    inv_s = torch.exp(net.variance * 10.0).clip(1e-6, 1e6)
    rot = torch.inverse(sc.c2ws[0, :3, :3])
    env, _ = composite_envelope(O, ref, sdf_g, grad_g, col_g, rays_o, rays_d, inv_s, rot, per_sample=per_sample)
    worst = {}
    for k in RAY_KEYS:
        if k == "val_normal" and "val_normal" not in out:
            continue
        refk = ref["_val_normal"] if k == "val_normal" else ref[k]
        worst[k] = assert_within_envelope(out[k], refk, env[k], what + k, rows=keep)
    if per_sample:
        worst["alpha"] = assert_within_envelope(out["_alpha"], ref["_alpha"], env["alpha"], what + "alpha", floor=1.0,
                                                rows=keep)
        worst["weights"] = assert_within_envelope(out["weights"], ref["weights"], env["weights"], what + "weights",
                                                  rows=keep)
        worst["weight_max"] = assert_within_envelope(out["weight_max"], ref["weight_max"],
                                                     env["weights"].max(dim=1, keepdim=True)[0], what + "weight_max",
                                                     rows=keep)
    return worst


def explain_view_mismatches(O, point_views, ref, sc):
    """Per-view validity census with proof: a (point, view) validity bit may differ from the oracle only for a
    projection within 1e-3 px of an image border / the w = 0 plane (projector.py:536; e.g. every ray of pixel row 0
    projects to y = 0 exactly in a source view that differs from the reference view only by an x offset).
    Returns the bool (B,) mask of rays WITHOUT such a sample."""
    B, S = ref["mid_z_vals"].shape
    vm = ref["_view_mask"]
    bits = torch.as_tensor(point_views).cpu()
    got = torch.stack([(bits >> v) & 1 for v in range(vm.shape[1])], dim=1).bool()
    mism = (got != vm).any(dim=1) & ref["_compute_mask"]
    if bool(mism.any()):
        border = O.projection_border_distance(ref["_pts"][mism], sc.intrs, sc.c2ws, sc.features)
        assert bool((border < 1e-3).all()), "view-mask mismatch %.3e px away from any image border" % float(border.max())
    return ~mism.reshape(B, S).any(dim=1)
