"""Timeline of CTA 0 of k_sdf_tc1<GRAD> (library built with SURF_NVCC_EXTRA=-DTC_TRACE)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from surf_b200 import _lib, synthetic
import bench
_lib.set_mlp_mode(int(os.environ.get('MLP_MODE', '1')))
sc = synthetic.make_scene(3, 576, 800, 88, seed=1, device="cuda")
m = bench.build_net("cuda")
ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
lib = _lib.load()
rays_o, rays_d, hw = synthetic.image_rays(sc, 1)
sel = slice(200 * 800, 200 * 800 + 8192)          # 8192 rays of image row 200.. (coherent, like the bench)
torch.manual_seed(0)
t_rand = m.draw_chunk_randoms(8192)
def run():
    return m.render_image(ps, rays_o[sel], rays_d[sel], sc.near, sc.far, t_rand=t_rand)
for _ in range(2):
    run()
torch.cuda.synchronize()
buf = (C.c_longlong * 8192)()
reader = lib.surf_t1_trace_read if os.environ.get("MLP_MODE") == "3" else lib.surf_t2_trace_read
reader.restype = C.c_int
reader(buf, 4096)
run()
torch.cuda.synchronize()
n = reader(buf, 4096)
ev = sorted((buf[2 * i + 1], buf[2 * i]) for i in range(n))
t0 = ev[0][0]
def name(e):
    if e == 1: return "stage_done"
    if e == 99: return "tile_done"
    if 10 <= e < 16: return "epi  dfull fwd L%d" % (e - 10)
    if 16 <= e < 22: return "epi  dfull bwd L%d" % (5 - (e - 16))
    if 30 <= e < 36: return "epi  done  fwd L%d" % (e - 30)
    if 36 <= e < 42: return "epi  done  bwd L%d" % (5 - (e - 36))
    if 50 <= e < 70: return "iss  start  phase %d" % (e - 50)
    if 70 <= e < 90: return "iss  issued phase %d" % (e - 70)
    if 1000 <= e < 2000: return "iss   wait w_full  p%d c%d" % ((e - 1000) // 8, (e - 1000) % 8)
    if 2000 <= e < 3000: return "iss   got  w_full  p%d c%d" % ((e - 2000) // 8, (e - 2000) % 8)
    if 6000 <= e < 7000:
        k = e - 6000
        return "epi    p%d g%d %s" % (k // 16, (k % 16) // 4, ("loaded", "computed", "signalled")[k % 4])
    if 4000 <= e < 5000: return "iss   feat chunk issued p%d h%d" % ((e - 4000) // 8, (e - 4000) % 8)
    if 3000 <= e < 4000: return "iss   committed    p%d c%d" % ((e - 3000) // 8, (e - 3000) % 8)
    return str(e)
prev = t0
for t, e in ev[:1500]:
    print("%9d (+%5d)  [%s] %s" % (t - t0, t - prev, ("w0 ", "iss", "w15")[e // 100000], name(e % 100000)))
    prev = t
