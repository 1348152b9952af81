"""GPU diagnostic: FAST vs REFERENCE render kernels over the test sweep (same LUTs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api, synthetic
W, H = 96, 54
dims = dict(scattering_r_size=16, scattering_mu_size=64, scattering_mu_s_size=16, scattering_nu_size=4)
bF = fb.Builder(0); bR = fb.Builder(0, kernels=api.KERNELS_REFERENCE)
pend = fb.Atmosphere.build(bF, None, fb.Parameters(**dims)); torch.cuda.synchronize()
atm = pend.atmosphere()
rF, rR = fb.Renderer(bF), fb.Renderer(bR)
draws, extra = synthetic.camera_sweep(24, W, H)
for k in range(24):
    depth = synthetic.analytic_depth(extra[k][0], extra[k][1], W, H)
    cF, tF = rF.draw_host(atm, draws[k], depth); cR, tR = rR.draw_host(atm, draws[k], depth)
    peak = max(float(np.abs(cR).max()), 1e-3)
    ec = np.abs(cF - cR) / np.maximum(np.abs(cR), 1e-3 * peak)
    et = np.abs(tF - tR) / np.maximum(np.abs(tR), 1e-6)
    w = np.unravel_index(ec.argmax(), ec.shape)
    print(f"view {k:2d} alt {draws[k].camera_position[2]-6360:9.3f} km ground {float((depth>0).mean()):.2f} color err max {ec.max():.2e} (fast {cF[w]:.4e} ref {cR[w]:.4e} depth {depth[w[0],w[1]]:.3e}) transm err max {et.max():.2e} nan {np.isnan(cF).sum()}")
