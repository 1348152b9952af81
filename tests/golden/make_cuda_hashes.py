"""SHA-256 of what the CUDA product path writes for a fixed small scenario (tables at three dims, two rendered views),
committed as tests/golden/cuda_hashes.json.  NOT a parity reference (parity is against the oracle and the reference's
golden vectors): a change detector.  Every optimisation of round 2 that claimed "bit-identical" was checked by exactly
this kind of hash A/B (profiles/r2_packed_pairs_ab.txt); the committed hashes make the claim checkable by
tests/test_parity_gpu.py::test_product_outputs_are_bit_stable for whatever comes next.  A change that is meant to alter
bits (within the tolerance) regenerates the file on a B200:  python tests/golden/make_cuda_hashes.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "smoke": dict(scattering_r_size=8, scattering_mu_size=32, scattering_mu_s_size=8, scattering_nu_size=2),
    "odd": dict(scattering_r_size=7, scattering_mu_size=22, scattering_mu_s_size=11, scattering_nu_size=3,
                transmittance_mu_size=100, transmittance_r_size=33, irradiance_mu_s_size=20, irradiance_r_size=9),
    "wide": dict(scattering_r_size=4, scattering_mu_size=8, scattering_mu_s_size=64, scattering_nu_size=32, order=3),
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute():
    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import synthetic
    b = fb.Builder(0)
    out = {}
    for name, dims in CASES.items():
        T, S, E = fb.precompute_host(b, fb.Parameters(**dims))
        out[name] = {"transmittance": digest(T), "scattering": digest(S), "irradiance": digest(E)}
    import torch
    pend = fb.Atmosphere.build(b, None, fb.Parameters(**CASES["smoke"]))
    torch.cuda.synchronize()
    atm = pend.assert_ready()
    r = fb.Renderer(b)
    W, H = 160, 90
    draws, extra = synthetic.camera_sweep(24, W, H)
    for k in (3, 11):
        depth = synthetic.analytic_depth(extra[k][0], extra[k][1], W, H)
        c, t = r.draw_host(atm, draws[k], depth)
        out[f"view{k}"] = {"color": digest(c), "transmittance": digest(t)}
    return out


if __name__ == "__main__":
    h = compute()
    with open(os.path.join(HERE, "cuda_hashes.json"), "w") as f:
        json.dump(h, f, indent=1, sort_keys=True)
    print(json.dumps(h, indent=1, sort_keys=True))
