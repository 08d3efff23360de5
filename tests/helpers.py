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


def blend_envelope(O, net, fv, rd, mk, n_random=4, seed=0):
    """Conditioning envelope of the reference's anti-alias pooling weights.

    weight = (exp(s (dot_v - 1)) - min_v exp(...)) / (sum + 1e-8) subtracts nearly equal fp32 exponentials
    (blending_network.py:76-80).  With 3+ source views of similar viewing angle a ONE-ulp change of one
    exponential moves the output by up to 0.2 in RGB (measured on the 5-view golden case), and torch's own
    CPU exp is not correctly rounded (differs from the correctly rounded value in ~2 % of arguments), so
    no independent implementation can match the reference bit-for-bit there.  Returns (reference output,
    per-point envelope) where the envelope is the largest change of the oracle's output when the
    exponentials move by +-1 ulp (each view alone, both signs, plus random patterns)."""
    g = torch.Generator().manual_seed(seed)
    base = O.blend(net, fv, rd, mk)
    n, V = fv.shape[0], fv.shape[1]
    env = torch.zeros(n)
    pats = []
    for v in range(V):
        for sgn in (-1, 1):
            p = torch.zeros(n, V, dtype=torch.int64)
            p[:, v] = sgn
            pats.append(p)
    for _ in range(n_random):
        pats.append(torch.randint(0, 3, (n, V), generator=g) - 1)
    for p in pats:
        env = torch.maximum(env, (O.blend(net, fv, rd, mk, e_ulp=p) - base).abs().max(dim=1)[0])
    return base, env
