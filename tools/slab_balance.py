"""Per-r-slab device time of every 3-D stage at the high-resolution dims (BASELINE.json configs[2]): how uneven is an
even split of the r axis over 8 ranks?  One GPU; stages run on whatever the images hold (timing only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dims = dict(transmittance_mu_size=1024, transmittance_r_size=256, scattering_r_size=128 // scale, scattering_mu_size=512 // scale,
            scattering_mu_s_size=128 // scale, scattering_nu_size=32 // scale)
b = fb.Builder(0)
s = torch.cuda.Stream()
p = fb.Atmosphere.build(b, s, fb.Parameters(order=3, **dims))
s.synchronize()
R, W = dims["scattering_r_size"], 8
m = R // W
rows = []
with torch.cuda.stream(s):
    for name, st, order in (("single", api.STAGE_SINGLE_SCATTERING, 0), ("density2", api.STAGE_SCATTERING_DENSITY, 2),
                            ("density3", api.STAGE_SCATTERING_DENSITY, 3), ("multiple", api.STAGE_MULTIPLE_SCATTERING, 0)):
        t = []
        for k in range(W):
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s); p.run_stage(st, order=order, r_begin=k * m, r_end=(k + 1) * m, stream=s); e1.record(s); e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            t.append(best)
        t = np.array(t)
        print(f"{name}: slab ms " + " ".join(f"{v:.2f}" for v in t) + f" | max/mean {t.max() / t.mean():.3f}")
        rows.append(t)
tot = np.sum(rows, axis=0) + np.array(rows[2]) * 5 + np.array(rows[3]) * 6      # 8 orders: density2 + 6 x density3, 7 x multiple
print("8-order total per slab ms: " + " ".join(f"{v:.1f}" for v in tot) + f" | max/mean {tot.max() / tot.mean():.3f}")
pair = tot.reshape(-1)  # mirrored pairing over 16 half-slabs is estimated from 8 slabs: pair (k, 7-k) halves
print("mirror-paired estimate max/mean: %.3f" % (max((tot[k] + tot[7 - k]) / 2 for k in range(4)) / tot.mean()))
