"""Profiling driver: a few 4K draws on default-dims LUTs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import synthetic
W, H = 3840, 2160
b = fb.Builder(0)
pend = fb.Atmosphere.build(b, None, fb.Parameters()); torch.cuda.synchronize()
atm = pend.atmosphere()
r = fb.Renderer(b)
draws, extra = synthetic.camera_sweep(16, W, H)
views = [1, 2, 11, 13]   # ground-heavy, space, low altitude, mixed
depth = torch.stack([torch.from_numpy(synthetic.analytic_depth(extra[k][0], extra[k][1], W, H)) for k in views]).cuda()
color = torch.empty((len(views), H, W, 4), device="cuda"); transm = torch.empty_like(color)
for rep in range(2):
    for i, k in enumerate(views):
        r.set_depth_buffer(0, depth[i])
        r.draw(None, atm, 0, draws[k], color[i], transm[i], W, H)
torch.cuda.synchronize()
print("ground fraction per view:", [(depth[i] > 0).float().mean().item() for i in range(len(views))])
