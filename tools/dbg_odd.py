import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api
dims = dict(transmittance_mu_size=37, transmittance_r_size=11, irradiance_mu_s_size=13, irradiance_r_size=5,
            scattering_r_size=3, scattering_mu_size=10, scattering_mu_s_size=6, scattering_nu_size=2)
for rep in range(4):
    b = fb.Builder(0)
    p = fb.Atmosphere.allocate(b, fb.Parameters(order=3, **dims))
    seq = [(0,0),(1,0),(2,0),(6,0),(3,2),(4,1),(5,0),(3,3),(4,2),(5,0)]
    for st, o in seq:
        try:
            p.run_stage(st, order=o); torch.cuda.synchronize()
        except Exception as e:
            print("rep", rep, "stage", st, "order", o, "->", e); break
    else:
        print("rep", rep, "ok")
    try:
        T,S,E = fb.precompute_host(b, fb.Parameters(order=3, **dims)); print("rep", rep, "host ok")
    except Exception as e:
        print("rep", rep, "host ->", e)
