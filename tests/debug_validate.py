"""ad-hoc GPU diagnostic (not a test): per-ray differences of validate() vs the golden image."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import surf_oracle as O
from helpers import load_golden, scene_from_recipe
from surf_b200 import conf
from surf_b200.modules.implicit_surface import ImplicitSurface
g = load_golden("validate_24x32"); sc = scene_from_recipe(g["recipe"])
m = ImplicitSurface(conf.default_implicit_surface_conf()); m.load_state_dict(g["sd"]); m = m.cuda()
d = sc.to("cuda"); i = g["in"]
net = O.OracleNet(g["sd"])
n = i["rays_o"].shape[0]
torch.manual_seed(0)
t_rand = m.draw_chunk_randoms(n)
ps = m.prepare(d.matching_volume, d.volumes, d.sparse_idxes, d.mask_volumes, d.imgs, d.features, d.intrs, d.c2ws)
near = i["near"]; far = i["far"]
# oracle per chunk with stages
bad_total = 0
for c in range(0, n, 256):
    sl = slice(c, min(n, c + 256))
    ref = O.render(net, i["rays_o"][sl], i["rays_d"][sl], near[sl], far[sl], sc.matching_volume, sc.volumes, sc.sparse_idxes,
                   sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, 1.0, t_rand=t_rand[sl],
                   pts_random=torch.zeros(4, 3), return_stages=True)
    out = m.render(i["rays_o"][sl].cuda(), i["rays_d"][sl].cuda(), near[sl].cuda(), far[sl].cuda(), ps, None, None, None,
                   None, None, None, None, None, 1.0, None, t_rand=t_rand[sl], pts_random=torch.zeros(4, 3),
                   return_stages=True)
    vm = (out["_point_flags"].cpu() & 1).bool()
    mm = (vm != ref["_voxel_mask"]).reshape(-1, 136).sum(1)
    ce = (out["color_fine"].cpu() - ref["color_fine"]).abs().max(1)[0]
    de = (out["sdf_depth"].cpu() - ref["sdf_depth"]).abs().max(1)[0]
    re = (out["render_depth"].cpu() - ref["render_depth"]).abs()
    ze = (out["mid_z_vals"].cpu() - ref["mid_z_vals"]).abs().max(1)[0]
    bad = (ce > 5e-4) | (de > 5e-4 * 3) | (re > 5e-4 * 3)
    print("chunk", c, "bad rays", int(bad.sum()), "mask-mismatch rays", int((mm > 0).sum()), "max zerr %.2e" % float(ze.max()))
    for r in torch.nonzero(bad)[:, 0].tolist():
        pi = (out["_prev_idx"].cpu()[r].item(), ref["_prev_idx"][r].item())
        print("   ray", c + r, "col %.2e sdfd %.2e rend %.2e maskmm %d zerr %.2e prev_idx %s sdf_depth %.4f/%.4f wsum %.4f/%.4f" % (
            ce[r], de[r], re[r], mm[r], ze[r], pi, out["sdf_depth"].cpu()[r], ref["sdf_depth"][r], out["weight_sum"].cpu()[r], ref["weight_sum"][r]))
        sd_o = out["sparse_sdf"].cpu()[4:].reshape(-1, 136)[r]; sd_r = ref["sparse_sdf"][4:].reshape(-1, 136)[r]
        print("      max sdf err %.2e  max alpha err %.2e  grad err %.2e" % (float((sd_o - sd_r).abs().max()),
              float((out["_alpha"].cpu()[r] - ref["_alpha"][r]).abs().max()), float((out["gradients"].cpu()[r] - ref["gradients"][r]).abs().max())))

print("---- validate() keys for ray 435")
torch.manual_seed(0)
hw = (24, 32)
out = m.validate(i["rays_o"].cuda(), i["rays_d"].cuda(), i["near"].cuda(), i["far"].cuda(), ps, None, None, None, None, None, None,
                 d.intrs, d.c2ws, None, None, hw, 1.0, None, extract_geometry=False)
for k in ["color_fine", "img_fine", "normal_img", "sdf_depth", "render_depth"]:
    a = torch.as_tensor(np.asarray(out[k])).double().reshape(768, -1)
    b = torch.as_tensor(np.asarray(g["out"][k])).double().reshape(768, -1)
    print(k, "ray435 err", (a[435] - b[435]).abs().max().item(), "a", a[435].tolist(), "b", b[435].tolist(), "scale", float(b.abs().max()))

print("---- ray 435 per-sample")
sl = slice(256, 512); r = 435 - 256
ref = O.render(net, i["rays_o"][sl], i["rays_d"][sl], near[sl], far[sl], sc.matching_volume, sc.volumes, sc.sparse_idxes,
               sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws, 1.0, t_rand=t_rand[sl], pts_random=torch.zeros(4, 3), return_stages=True)
o2 = m.render(i["rays_o"][sl].cuda(), i["rays_d"][sl].cuda(), near[sl].cuda(), far[sl].cuda(), ps, None, None, None,
              None, None, None, None, None, 1.0, None, t_rand=t_rand[sl], pts_random=torch.zeros(4, 3), return_stages=True)
w_o, w_r = o2["weights"].cpu()[r], ref["weights"][r]
print("weights err", float((w_o - w_r).abs().max()), "wsum", float(w_o.sum()), float(w_r.sum()))
a_o, a_r = o2["_alpha"].cpu()[r], ref["_alpha"][r]
j = int((a_o - a_r).abs().argmax())
print("alpha err max at", j, float(a_o[j]), float(a_r[j]))
sd_o = o2["sparse_sdf"].cpu()[4:].reshape(-1, 136)[r]; sd_r = ref["sparse_sdf"][4:].reshape(-1, 136)[r]
print("sdf at j", float(sd_o[j]), float(sd_r[j]), "grad", o2["gradients"].cpu()[r, j].tolist(), ref["gradients"][r, j].tolist())
nz = torch.nonzero(w_r > 1e-4)[:, 0].tolist()
print("samples with weight:", nz)
for jj in nz[:12]:
    print(jj, "w %.5f/%.5f alpha %.5f/%.5f sdf %.6f/%.6f" % (w_o[jj], w_r[jj], a_o[jj], a_r[jj], sd_o[jj], sd_r[jj]), "g", [round(x, 4) for x in o2["gradients"].cpu()[r, jj].tolist()], [round(x, 4) for x in ref["gradients"][r, jj].tolist()])
vn_r = (ref["gradients"] * ref["weights"][:, :, None] * ref["inside_sphere"][..., None]).sum(1)[r]
print("val_normal gpu(render_image) vs oracle:", vn_r.tolist())
