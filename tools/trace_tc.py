"""Builds sdf_tc.cu with -DTC_TRACE into a side library and prints the event timeline of CTA 0."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from surf_b200 import _lib, synthetic, build as B
import bench
_lib.set_mlp_mode(int(os.environ.get('MLP_MODE', '1')))
sc = synthetic.make_scene(3, 576, 800, 88, seed=1, device="cuda")
m = bench.build_net("cuda")
ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
lib = _lib.load()
for _ in range(2):
    u = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], 512, x_range=(200, 216))
torch.cuda.synchronize()
buf = (C.c_longlong * 4096)()
lib.surf_tc_trace_read.restype = C.c_int
n = lib.surf_tc_trace_read(buf, 2048)      # discard warm-up
u = m.sdf_grid(ps, [-1, -1, -1], [1, 1, 1], 512, x_range=(200, 216))
torch.cuda.synchronize()
n = lib.surf_tc_trace_read(buf, 2048)
ev = sorted((buf[2 * i + 1], buf[2 * i]) for i in range(n))
t0 = ev[0][0]
names = {1: "stage_done", 10: "dfull", 20: "epi_done", 30: "iss_aready", 40: "iss_issued"}
for t, e in ev[:220]:
    g, k = divmod(e, 100)
    base = 1 if k == 1 else (k // 10) * 10
    print("%9d  tile%s %-11s L%d" % (t - t0, "XY"[g], names[base], k - base if base != 1 else 0))
