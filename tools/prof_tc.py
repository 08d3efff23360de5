import sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))); 
from surf_b200 import _lib, synthetic
import bench
_lib.set_mlp_mode(int(__import__('os').environ.get('MLP_MODE', '1')))
sc = synthetic.make_scene(3, 576, 800, 88, seed=1, device="cuda")
m = bench.build_net("cuda")
ps = m.prepare(sc.matching_volume, sc.volumes, sc.sparse_idxes, sc.mask_volumes, sc.imgs, sc.features, sc.intrs, sc.c2ws)
for _ in range(2):
    u = m.sdf_grid(ps, [-1,-1,-1],[1,1,1], 512, x_range=(200, 232))
torch.cuda.synchronize()
import time
t0=time.time()
for _ in range(5):
    u = m.sdf_grid(ps, [-1,-1,-1],[1,1,1], 512, x_range=(200, 232))
torch.cuda.synchronize()
print('ms per 8.39M pts:', (time.time()-t0)/5*1e3, float(u.mean()))
