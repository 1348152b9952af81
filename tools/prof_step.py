"""Profiling driver: one warm-up build + N resubmits of the default-dims precompute (no timing claims)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuzzyblue_b200 as fb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
b = fb.Builder(0)
s = torch.cuda.Stream()
p = fb.Atmosphere.build(b, s, fb.Parameters())
s.synchronize()
for _ in range(n):
    # stage by stage (not the graph) so that ncu sees plain kernel launches
    from fuzzyblue_b200 import api
    p.run_stage(api.STAGE_TRANSMITTANCE, stream=s); p.run_stage(api.STAGE_DIRECT_IRRADIANCE, stream=s)
    p.run_stage(api.STAGE_SINGLE_SCATTERING, stream=s); p.run_stage(api.STAGE_CLEAR_IRRADIANCE, stream=s)
    for order in (2, 3, 4):
        p.run_stage(api.STAGE_SCATTERING_DENSITY, order=order, stream=s)
        p.run_stage(api.STAGE_INDIRECT_IRRADIANCE, order=order - 1, stream=s)
        p.run_stage(api.STAGE_MULTIPLE_SCATTERING, stream=s)
s.synchronize()
