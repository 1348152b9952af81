"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): reduced, odd and wide dims, both orders' density
variants, the renderer and the library queries."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api, synthetic
b = fb.Builder(0)
cases = [dict(scattering_r_size=4, scattering_mu_size=8, scattering_mu_s_size=8, scattering_nu_size=2, order=3),
         dict(scattering_r_size=3, scattering_mu_size=4, scattering_mu_s_size=12, scattering_nu_size=6, order=3),     # nu not a power of two
         dict(scattering_r_size=2, scattering_mu_size=4, scattering_mu_s_size=7, scattering_nu_size=3, order=3),
         dict(scattering_r_size=2, scattering_mu_size=4, scattering_mu_s_size=64, scattering_nu_size=32, order=3),
         dict(scattering_r_size=3, scattering_mu_size=6, scattering_mu_s_size=5, scattering_nu_size=4, order=3,
              transmittance_mu_size=37, transmittance_r_size=11, irradiance_mu_s_size=13, irradiance_r_size=5)]
for d in cases:
    T, S, E = fb.precompute_host(b, fb.Parameters(**d))
    assert np.isfinite(S.astype(np.float32)).all() and np.isfinite(E).all()
p = fb.Parameters(**cases[0])
pend = fb.Atmosphere.build(b, None, p); torch.cuda.synchronize()
pend.resubmit(None); torch.cuda.synchronize()
# read-backs recorded into the command stream (r-slabs of the last multiple-scattering pass + their copies)
hT = torch.zeros((p.transmittance_r_size, p.transmittance_mu_size, 4)).pin_memory()
hE = torch.zeros((p.irradiance_r_size, p.irradiance_mu_s_size, 4)).pin_memory()
hS = torch.zeros((p.scattering_r_size, p.scattering_mu_size, p.scattering_nu_size * p.scattering_mu_s_size, 4), dtype=torch.float16).pin_memory()
pend.set_readback(hT.data_ptr(), hS.data_ptr(), hE.data_ptr())
pend.resubmit(None); torch.cuda.synchronize()
assert np.array_equal(hS.numpy().view(np.uint16), pend.atmosphere().read_scattering().view(np.uint16))
pend.set_readback(None, None, None)
atm = pend.assert_ready()
r = fb.Renderer(b)
draws, extra = synthetic.camera_sweep(14, 64, 36)
for k in (1, 2, 13):
    c, t = r.draw_host(atm, draws[k], synthetic.analytic_depth(extra[k][0], extra[k][1], 64, 36))
    assert np.isfinite(c).all()
# the sharded schedule at world = 1 (row-ranged indirect irradiance, slab stages) and the stream-ordered block cache
from fuzzyblue_b200 import sharded
sp = sharded.build_sharded(b, p, None, 0, 1); sp.wait()
sp.run_stage(api.STAGE_INDIRECT_IRRADIANCE, order=2, r_begin=1, r_end=3); sp.wait()
early = fb.Atmosphere.build(b, torch.cuda.Stream(), p); early.resubmit(None)
early.assert_ready(check=False).close()
again = fb.Atmosphere.build(b, None, p); again.wait(); again.assert_ready().close()
print("sanitize pass done")
