"""Error of the tensor-core SDF / gradient kernels against the committed reference goldens, per mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_golden, scene_from_recipe
from surf_b200 import _lib, conf
from surf_b200.modules.implicit_surface import ImplicitSurface
for name in ("render_v2_perturbed", "render_v2_init", "render_v4_perturbed"):
    g = load_golden(name)
    d = scene_from_recipe(g["recipe"]).to("cuda")
    m = ImplicitSurface(conf.default_implicit_surface_conf()); m.load_state_dict(g["sd"]); m = m.cuda()
    ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
    pv = torch.from_numpy(g["out"]["_pts_valid"]).cuda()
    rs = torch.from_numpy(g["out"]["_sdf_full"][:, :1]).double()
    rg = torch.from_numpy(g["out"]["_grad_valid"]).double()
    for mode in (0, 3, 1):
        _lib.set_mlp_mode(mode)
        s, gr = m.sdf_network.gradient(pv, ps, with_sdf=True)
        es = (s.cpu().double() - rs).abs(); eg = (gr.cpu().double() - rg).abs()
        print("%-20s mode %d  sdf: max %.2e rms %.2e (scale %.2e)   grad: max %.2e rms %.2e (scale %.2e)" % (
            name, mode, es.max(), es.pow(2).mean().sqrt(), rs.abs().max(), eg.max(), eg.pow(2).mean().sqrt(), rg.abs().max()))
_lib.set_mlp_mode(0)
