"""Per-stage device times for a few random atmospheres (config-4 physics); FUZZYBLUE_B200_LIB selects a variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api, synthetic
b = fb.Builder(0)
s = torch.cuda.Stream()
for k, prm in enumerate(synthetic.random_atmospheres(6, seed=20260)):
    p = fb.Atmosphere.build(b, s, prm); s.synchronize()
    plan = [("single", api.STAGE_SINGLE_SCATTERING, 0), ("density2", api.STAGE_SCATTERING_DENSITY, 2), ("multiple", api.STAGE_MULTIPLE_SCATTERING, 0),
            ("density3", api.STAGE_SCATTERING_DENSITY, 3)]
    out = []
    with torch.cuda.stream(s):
        for name, st, order in plan:
            ts = []
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s); p.run_stage(st, order=order, stream=s); e1.record(s); e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            out.append(f"{name} {np.median(ts[1:]):.0f}")
    print(k, "bottom %.0f top %.0f" % (prm.bottom_radius, prm.top_radius), ", ".join(out))
    p.close()
