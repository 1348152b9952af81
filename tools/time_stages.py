"""Per-stage device times (CUDA events, median of N) of the default-dims precompute; FUZZYBLUE_B200_LIB selects a variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
b = fb.Builder(0)
s = torch.cuda.Stream()
p = fb.Atmosphere.build(b, s, fb.Parameters())
s.synchronize()
plan = [("transmittance", api.STAGE_TRANSMITTANCE, 0), ("direct_irr", api.STAGE_DIRECT_IRRADIANCE, 0), ("single", api.STAGE_SINGLE_SCATTERING, 0),
        ("clear", api.STAGE_CLEAR_IRRADIANCE, 0)]
for order in (2, 3, 4):
    plan += [(f"density{order}", api.STAGE_SCATTERING_DENSITY, order), (f"indirect{order-1}", api.STAGE_INDIRECT_IRRADIANCE, order - 1),
             (f"multiple{order}", api.STAGE_MULTIPLE_SCATTERING, 0)]
times = {k: [] for k, _, _ in plan}
with torch.cuda.stream(s):
    for rep in range(n + 2):
        for name, st, order in plan:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); p.run_stage(st, order=order, stream=s); e1.record(s); e1.synchronize()
            if rep >= 2: times[name].append(e0.elapsed_time(e1) * 1e3)
tot = 0.0
out = []
for name, _, _ in plan:
    m = float(np.median(times[name])); tot += m; out.append(f"{name} {m:.0f}")
print(os.environ.get("FUZZYBLUE_B200_LIB", "default"), "| total %.0f us |" % tot, ", ".join(out))
