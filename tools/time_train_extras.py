import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import torch, bench
from surf_b200 import synthetic
sc = synthetic.make_scene(5, 480, 640, 64, seed=10, device="cuda")
m = bench.build_net("cuda")
ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
g = torch.Generator(device="cuda").manual_seed(0)
pts = torch.nn.functional.normalize(torch.randn(69632, 3, device="cuda", generator=g), dim=1) * 0.5
fl = torch.full((69632,), 2, dtype=torch.uint8, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("smooth kernel (tcgen05), 69632 points: %.3f ms" % t(lambda: m.sdf_network.smooth(pts, ps, flags=fl, mode=1)))
print("smooth kernel (fp32 FFMA), 69632 points: %.3f ms" % t(lambda: m.sdf_network.smooth(pts, ps, flags=fl, mode=0)))
print("gradient (tcgen05), 69632 points: %.3f ms" % t(lambda: m.sdf_network.gradient(pts, ps, with_sdf=True)))
o, d = synthetic.random_pixel_rays(sc, 512, seed=1)
o, d = o.cuda(), d.cuda()
near, far = sc.near.expand(512, 1).contiguous(), sc.far.expand(512, 1).contiguous()
print("render() 512 rays with extras: %.3f ms" % t(lambda: m.render(o, d, near, far, ps, None, None, None, None, None, None, sc.intrs, sc.c2ws, 1.0, None)))
print("render() 512 rays without extras: %.3f ms" % t(lambda: m.render(o, d, near, far, ps, None, None, None, None, None, None, sc.intrs, sc.c2ws, 1.0, None, extras=False)))
t0 = time.perf_counter()
for _ in range(5):
    m.render(o, d, near, far, ps, None, None, None, None, None, None, sc.intrs, sc.c2ws, 1.0, None)
torch.cuda.synchronize()
print("wall per render(): %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
