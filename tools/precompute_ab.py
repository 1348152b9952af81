"""A/B of two builds of the precompute: runs the default-dims 4-order precompute and a reduced / odd-dims one with the
library FUZZYBLUE_B200_LIB selects and writes one SHA-256 per table (transmittance, scattering, irradiance and the
last order's temporaries) to argv[1]; run it once per build and diff the files.  Bit-identical variants must produce
identical files.  Also prints per-stage device times (median of argv[2] runs, default 10)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api

out = open(sys.argv[1], "w")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
b = fb.Builder(0)
s = torch.cuda.Stream()

def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]

IMAGES = [("transmittance", api.IMAGE_TRANSMITTANCE), ("irradiance", api.IMAGE_IRRADIANCE), ("scattering", api.IMAGE_SCATTERING),
          ("delta_irradiance", api.IMAGE_DELTA_IRRADIANCE), ("scattering_density", api.IMAGE_SCATTERING_DENSITY),
          ("delta_multiple_scattering", api.IMAGE_DELTA_MULTIPLE_SCATTERING)]
CASES = [("default", dict()),
         ("reduced", dict(scattering_r_size=16, scattering_mu_size=64, scattering_mu_s_size=16, scattering_nu_size=4)),
         ("wide", dict(scattering_r_size=4, scattering_mu_size=8, scattering_mu_s_size=64, scattering_nu_size=32)),   # tests/conftest.py WIDE_DIMS: W = 2048, 1024 threads x 2 texels
         ("odd", dict(scattering_r_size=7, scattering_mu_size=22, scattering_mu_s_size=11, scattering_nu_size=3,
                      transmittance_mu_size=100, transmittance_r_size=33, irradiance_mu_s_size=20, irradiance_r_size=9))]
pend_default = None
for name, dims in CASES:
    p = fb.Atmosphere.build(b, s, fb.Parameters(**dims))
    s.synchronize()
    for iname, image in IMAGES:
        a = p.download(image, s)
        s.synchronize()
        print(name, iname, digest(a), file=out)
    if name == "default":
        pend_default = p
out.close()

plan = [("transmittance", api.STAGE_TRANSMITTANCE, 0), ("single", api.STAGE_SINGLE_SCATTERING, 0)]
for order in (2, 3):
    plan += [(f"density{order}", api.STAGE_SCATTERING_DENSITY, order), (f"indirect{order - 1}", api.STAGE_INDIRECT_IRRADIANCE, order - 1),
             (f"multiple{order}", api.STAGE_MULTIPLE_SCATTERING, 0)]
times = {k: [] for k, _, _ in plan}
with torch.cuda.stream(s):
    for rep in range(n + 2):
        for name, st, order in plan:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); pend_default.run_stage(st, order=order, stream=s); e1.record(s); e1.synchronize()
            if rep >= 2:
                times[name].append(e0.elapsed_time(e1) * 1e3)
print("stage us (median of %d): " % n + ", ".join("%s %.0f" % (k, float(np.median(v))) for k, v in times.items()))
