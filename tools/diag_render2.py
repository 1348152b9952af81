"""GPU diagnostic: FAST vs REFERENCE render kernels, sweep views + cameras around the top boundary (same LUTs)."""
import sys, os, copy
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
alts = (59.0, 59.67, 59.69, 59.99, 60.0, 60.001, 61.0, 1e-4, 3.0)
views = []
for seed, altitudes, tag in ((11, None, "sweep"), (12, alts, "fixed")):
    draws, extra = synthetic.camera_sweep(16 if altitudes is None else len(alts), W, H, seed=seed, altitudes_km=altitudes)
    views += [(d, synthetic.analytic_depth(inv, eye, W, H), tag) for d, (inv, eye) in zip(draws, extra)]
for k, (d, depth, tag) in enumerate(views):
    cF, tF = rF.draw_host(atm, d, depth); cR, tR = rR.draw_host(atm, d, depth)
    ok = np.isfinite(cR).all(axis=-1) & np.isfinite(tR).all(axis=-1)
    peak = max(float(np.abs(cR[ok]).max()), 1e-3)
    ec = np.where(ok[..., None], np.abs(cF - cR) / np.maximum(np.abs(cR), 1e-3 * peak), 0)
    et = np.where(ok[..., None], np.abs(tF - tR) / np.maximum(np.abs(tR), 1e-6), 0)
    w = np.unravel_index(ec.argmax(), ec.shape); wt = np.unravel_index(et.argmax(), et.shape)
    print(f"{tag} {k:2d} alt {d.camera_position[2]-6360:9.4f} ground {float((depth>0).mean()):.2f} bad {int((~ok).sum())} "
          f"color {ec.max():.2e} (F {cF[w]:.5e} R {cR[w]:.5e} depth {depth[w[0],w[1]]:.3e}) transm {et.max():.2e} (F {tF[wt]:.5e} R {tR[wt]:.5e}) finite-eq {np.array_equal(np.isfinite(cF), np.isfinite(cR))}")
