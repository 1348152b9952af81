"""Generates the committed fixtures under tests/golden/ from the CPU oracle (python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot be built here (no cargo / shaderc / Vulkan ICD), so these are
snapshots of oracle/fb_oracle.cpp — PARITY UNPINNED beyond the closed-form KATs of tests/test_oracle_kat.py:
  smoke_f32.npz     every table of a 4-order precompute at the dims of the reference's tests/smoke.rs:136-142, mode fp32
  default_f32.npz   default dims (Parameters::default(), 4 orders): transmittance + irradiance in full, and 4096 seeded
                    texels of every 3-D table of every order, mode fp32 (the parity target), plus the fp64-ideal values of
                    the same texels of the final scattering table (information only)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SMOKE_DIMS = dict(scattering_r_size=8, scattering_mu_size=32, scattering_mu_s_size=8, scattering_nu_size=2)


def main():
    t = time.time()
    s = O.precompute(O.Params(**SMOKE_DIMS), O.F32, keep_history=True)
    out = dict(transmittance=s.transmittance.astype(np.float32), irradiance=s.irradiance.astype(np.float32),
               scattering=s.scattering.astype(np.float16), delta_rayleigh=s.delta_rayleigh.astype(np.float16),
               delta_mie=s.delta_mie.astype(np.float16))
    for order in (2, 3, 4):
        h = s.history[order]
        out[f"o{order}_scattering_density"] = h["scattering_density"].astype(np.float16)
        out[f"o{order}_delta_multiple_scattering"] = h["delta_multiple_scattering"].astype(np.float16)
        out[f"o{order}_delta_irradiance"] = h["delta_irradiance"].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "smoke_f32.npz"), **out)
    print("smoke dims done", time.time() - t)

    p = O.Params()
    d = O.precompute(p, O.F32, keep_history=True)
    print("default dims fp32 done", time.time() - t)
    n_tex = int(np.prod(p.s_shape[:3]))
    idx = np.sort(np.random.default_rng(1234).choice(n_tex, 4096, replace=False)).astype(np.int64)
    g = dict(idx=idx, transmittance=d.transmittance.astype(np.float32), irradiance=d.irradiance.astype(np.float32),
             scattering=d.scattering.reshape(-1, 4)[idx].astype(np.float16),
             delta_rayleigh=d.delta_rayleigh.reshape(-1, 4)[idx].astype(np.float16),
             delta_mie=d.delta_mie.reshape(-1, 4)[idx].astype(np.float16),
             direct_irradiance=d.history["single"]["delta_irradiance"].astype(np.float32))
    for order in (2, 3, 4):
        h = d.history[order]
        g[f"o{order}_scattering_density"] = h["scattering_density"].reshape(-1, 4)[idx].astype(np.float16)
        g[f"o{order}_delta_multiple_scattering"] = h["delta_multiple_scattering"].reshape(-1, 4)[idx].astype(np.float16)
        g[f"o{order}_scattering"] = h["scattering"].reshape(-1, 4)[idx].astype(np.float16)
        g[f"o{order}_delta_irradiance"] = h["delta_irradiance"].astype(np.float32)
        g[f"o{order}_irradiance"] = h["irradiance"].astype(np.float32)
    i = O.precompute(p, O.F64)
    print("default dims fp64 done", time.time() - t)
    g["scattering_f64_ideal"] = i.scattering.reshape(-1, 4)[idx]
    g["irradiance_f64_ideal"] = i.irradiance
    g["transmittance_f64_ideal"] = i.transmittance
    np.savez_compressed(os.path.join(HERE, "default_f32.npz"), **g)
    print("wrote fixtures", time.time() - t)


if __name__ == "__main__":
    main()
