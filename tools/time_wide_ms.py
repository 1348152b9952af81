"""Multiple / single scattering and density at the high-resolution dims (rows of 4096 texels): device time of a 16-level
slab and a hash of the outputs; FUZZYBLUE_B200_MS_WIDE_CH selects the nodes staged per pass."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api
dims = dict(transmittance_mu_size=1024, transmittance_r_size=256, scattering_r_size=128, scattering_mu_size=512,
            scattering_mu_s_size=128, scattering_nu_size=32)
b = fb.Builder(0)
s = torch.cuda.Stream()
p = fb.Atmosphere.allocate(b, fb.Parameters(order=3, **dims))
with torch.cuda.stream(s):
    p.run_stage(api.STAGE_TRANSMITTANCE, stream=s); p.run_stage(api.STAGE_DIRECT_IRRADIANCE, stream=s)
    p.run_stage(api.STAGE_SINGLE_SCATTERING, r_begin=0, r_end=16, stream=s); p.run_stage(api.STAGE_CLEAR_IRRADIANCE, stream=s)
    p.run_stage(api.STAGE_SCATTERING_DENSITY, order=2, r_begin=0, r_end=16, stream=s)
    out = {}
    for name, st, order in (("single", api.STAGE_SINGLE_SCATTERING, 0), ("density2", api.STAGE_SCATTERING_DENSITY, 2), ("multiple", api.STAGE_MULTIPLE_SCATTERING, 0),
                            ("density3", api.STAGE_SCATTERING_DENSITY, 3)):
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); p.run_stage(st, order=order, r_begin=0, r_end=16, stream=s); e1.record(s); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = best
s.synchronize()
def slab_hash(image):
    ptr, nbytes = p.image(image)
    class _Raw:
        __cuda_array_interface__ = {"shape": (nbytes // 8 // 8,), "typestr": "<u8", "data": (ptr, False), "version": 2}
    t = torch.as_tensor(_Raw(), device="cuda")      # first 1/8 of the table = the slab
    return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:16]
print("MS_WIDE_CH=%s: " % os.environ.get("FUZZYBLUE_B200_MS_WIDE_CH", "default") + ", ".join(f"{k} {v:.2f} ms" for k, v in out.items())
      + f" | dMS slab hash {slab_hash(api.IMAGE_DELTA_MULTIPLE_SCATTERING)} density(order 3) slab hash {slab_hash(api.IMAGE_SCATTERING_DENSITY)}")
