"""Profiling driver: one sharded-style pass (world 1) of the high-resolution workload at 1/scale per axis."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p = fb.Parameters(order=3, transmittance_mu_size=1024, transmittance_r_size=256, scattering_r_size=128 // scale,
                  scattering_mu_size=512 // scale, scattering_mu_s_size=128 // scale, scattering_nu_size=32 // scale)
b = fb.Builder(0)
pend = fb.Atmosphere.build(b, None, p)
torch.cuda.synchronize()
