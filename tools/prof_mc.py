"""GPU marching cubes on an analytic 512^3 grid (sphere + ripples), for ncu captures / timing of csrc/marching.cu."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from surf_b200 import mesh
n = int(os.environ.get("N", "512"))
g = torch.linspace(-1, 1, n, device="cuda")
X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
u = (0.5 - torch.sqrt(X * X + Y * Y + Z * Z) + 0.01 * torch.sin(40 * X) * torch.sin(37 * Y) * torch.sin(43 * Z)).contiguous()
del X, Y, Z
for _ in range(3):
    v, t = mesh.marching_cubes_device(u, 0.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    v, t = mesh.marching_cubes_device(u, 0.0)
e1.record()
torch.cuda.synchronize()
print("marching cubes %d^3: %.3f ms per call, %d vertices, %d triangles" % (n, e0.elapsed_time(e1) / 5, v.shape[0], t.shape[0]))
