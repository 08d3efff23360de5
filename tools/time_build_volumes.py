"""Timing of SuRF.init_volumes (pyramid + four-stage volume construction) at the benchmark size (576 x 800, 3 views,
volumes 88 -> 704) with a device-side stand-in for the torchsparse regularisation network."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
from surf_b200 import conf, synthetic
from surf_b200.surf import SuRF

DEV = "cuda:0"
base = int(sys.argv[1]) if len(sys.argv) > 1 else 88
c = conf.parse_string("""
    range_ratios = [1.0, 0.4, 0.1, 0.01]
    feature_network { d_in = 3, d_base = 8, d_out = [4, 4, 4, 4] }
    volume { base_volume_dim = [%d, %d, %d] }
    matching_field { n_samples_depths = [128, 64, 32, 16], n_importance_depths = [128, 64, 32, 16],
                     up_sample_steps = [4, 4, 4, 4], depth_res_levels = [4, 2, 2, 1] }
""" % (base, base, base))
c.put("implicit_surface", conf.default_implicit_surface_conf())
torch.manual_seed(0)
model = SuRF(c).to(DEV)
g = torch.Generator(device=DEV).manual_seed(1)
A = [torch.randn(i, 8, device=DEV, generator=g) * 0.5 for i in (8, 16, 16, 16)]
B = [torch.randn(i, 8, device=DEV, generator=g) * 0.5 for i in (8, 16, 16, 16)]


def reg(feats, coords, s):
    # logits peaked on the r = 0.5 sphere so that the depth filter keeps a shell like a real scene
    n = float(base * 2 ** s - 1)
    p = coords[:, 1:].float() / n * 2 - 1
    out = torch.tanh(feats @ A[s])
    out[:, 0] = -40.0 * (p.norm(dim=1) - 0.5).abs()
    return out, torch.tanh(feats @ B[s])


model.reg_network = reg
intrs, c2ws, near, far = synthetic.make_cameras(3, 576, 800, DEV)
imgs = torch.rand(3, 3, 576, 800, device=DEV)
near_fars = torch.stack([torch.tensor([float(near), float(far)])] * 3)
ipts = {"imgs": imgs, "intrs": intrs, "c2ws": c2ws, "near": near, "far": far, "near_fars": near_fars, "src_idx": 1}
for compact in (False, True):
    for it in range(2):
        torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats(); t0 = time.perf_counter()
        model.init_volumes(ipts, compact=compact)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if compact:
        st = model.prepared.stats()
        print("compact=True : %.1f ms, peak %.1f GB, scene %s" % (dt * 1e3, torch.cuda.max_memory_allocated() / 2 ** 30, st))
    else:
        print("compact=False: %.1f ms, peak %.1f GB, voxels %s" % (dt * 1e3, torch.cuda.max_memory_allocated() / 2 ** 30,
                                                                  [int(v.shape[0]) for v in model.volumes]))
