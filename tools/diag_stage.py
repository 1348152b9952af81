"""GPU diagnostic: run each FAST stage on the exact inputs of the REFERENCE family at given dims and report error stats."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api

dims = {}
if len(sys.argv) > 1 and sys.argv[1] == "dump":
    dims = dict(scattering_r_size=16, scattering_mu_size=64, scattering_mu_s_size=16, scattering_nu_size=4)
if len(sys.argv) > 1 and sys.argv[1] == "hires2":   # half-scale cut of BASELINE.json configs[2]: rows of 1024 texels
    dims = dict(scattering_r_size=16, scattering_mu_size=64, scattering_mu_s_size=64, scattering_nu_size=16,
                transmittance_mu_size=512, transmittance_r_size=128)
if len(sys.argv) > 1 and sys.argv[1] == "hires1":   # full-resolution rows (4096 texels), few r / mu
    dims = dict(scattering_r_size=8, scattering_mu_size=16, scattering_mu_s_size=128, scattering_nu_size=32,
                transmittance_mu_size=1024, transmittance_r_size=256)
p = fb.Parameters(**dims)
bR = fb.Builder(0, kernels=api.KERNELS_REFERENCE)
bF = fb.Builder(0, kernels=api.KERNELS_FAST)
R = fb.Atmosphere.allocate(bR, p)
Fp = fb.Atmosphere.allocate(bF, p)
IM3 = [api.IMAGE_SCATTERING, api.IMAGE_DELTA_RAYLEIGH, api.IMAGE_DELTA_MIE, api.IMAGE_SCATTERING_DENSITY, api.IMAGE_DELTA_MULTIPLE_SCATTERING]
ALL = [api.IMAGE_TRANSMITTANCE, api.IMAGE_IRRADIANCE, api.IMAGE_DELTA_IRRADIANCE] + IM3
names = {0: "T", 1: "E", 2: "S", 3: "dE", 4: "dR", 5: "dM", 6: "dens", 7: "dMS"}

def err(a, b, f16):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 2.0 ** -14 if f16 else 1e-30)

def step(stage, order, outs):
    # copy REFERENCE state into FAST pending, run the stage on both, compare outputs
    for im in ALL:
        Fp.upload(im, R.download(im))
    R.run_stage(stage, order=order)
    Fp.run_stage(stage, order=order)
    for im in outs:
        a, b = Fp.download(im), R.download(im)
        e = err(a, b, im in IM3)
        w = np.unravel_index(int(e.argmax()), e.shape)
        print(f"stage {stage} order {order} {names[im]:5s}: max {e.max():.3e} at {w} fast={a[w]} ref={b[w]}  >5e-4: {(e>5e-4).sum()}  >1e-3: {(e>1e-3).sum()} of {e.size}")

for im in ALL:
    R.upload(im, np.zeros(R._shape(im), dtype=np.float32 if im in (0, 1, 3) else np.float16))
step(api.STAGE_TRANSMITTANCE, 0, [api.IMAGE_TRANSMITTANCE])
step(api.STAGE_DIRECT_IRRADIANCE, 0, [api.IMAGE_DELTA_IRRADIANCE])
step(api.STAGE_SINGLE_SCATTERING, 0, [api.IMAGE_DELTA_RAYLEIGH, api.IMAGE_DELTA_MIE, api.IMAGE_SCATTERING])
for order in (2, 3, 4):
    step(api.STAGE_SCATTERING_DENSITY, order, [api.IMAGE_SCATTERING_DENSITY])
    step(api.STAGE_INDIRECT_IRRADIANCE, order - 1, [api.IMAGE_DELTA_IRRADIANCE, api.IMAGE_IRRADIANCE])
    step(api.STAGE_MULTIPLE_SCATTERING, 0, [api.IMAGE_DELTA_MULTIPLE_SCATTERING, api.IMAGE_SCATTERING])
