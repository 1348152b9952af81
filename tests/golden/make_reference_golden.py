"""Generates the golden vectors of tests/golden/reference_*.npz by RUNNING THE REFERENCE'S OWN SHADERS on the CPU
(oracle/ref_glsl.py: /root/reference/shaders/*.h, *.comp, render_sky.frag compiled as C++ by oracle/glsl_ref/).

    python tests/golden/make_reference_golden.py          # needs /root/reference (this container); ~2 min on 8 cores

  reference_smoke_f32.npz    every table of every order of a 4-order precompute at the dims of the reference's
                             tests/smoke.rs:136-142, plus three rendered 48x27 views on those tables
  reference_default_f32.npz  default dims (Parameters::default(), 4 orders; BASELINE.json configs[1]): transmittance,
                             irradiance and every delta_irradiance in full, 4096 seeded texels of every 3-D table of every
                             order (the sky evaluation is pinned at the smoke dims, where the full tables fit a fixture)
  reference_odd_f32.npz      dims that are neither powers of two nor multiples of a warp (nu 3, mu_s 11, mu 22, r 7,
                             transmittance 100x33, irradiance 20x9), 4 orders: every table of every order in full
  reference_wide_f32.npz     rows of 2048 texels with 32 nu knots (r 4, mu 8, mu_s 64, nu 32), 3 orders: the 2-D tables in
                             full, 4096 seeded texels of every 3-D table (the bank-swizzled density tables and the
                             multi-texel-per-thread row kernels of the high-resolution configuration run at these dims)
Values are stored in the reference's storage formats (float32 / float16), which hold them exactly.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fuzzyblue_b200 import synthetic   # noqa: E402  (host-side numpy only: camera sweep + analytic depth)
from oracle import oracle as O         # noqa: E402  (Params / pack_draw only)
from oracle import ref_glsl as R       # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SMOKE_DIMS = dict(scattering_r_size=8, scattering_mu_size=32, scattering_mu_s_size=8, scattering_nu_size=2)
VIEWS_SMALL = (3, 11, 13)


ODD_DIMS = dict(scattering_r_size=7, scattering_mu_size=22, scattering_mu_s_size=11, scattering_nu_size=3,
                transmittance_mu_size=100, transmittance_r_size=33, irradiance_mu_s_size=20, irradiance_r_size=9)
WIDE_DIMS = dict(scattering_r_size=4, scattering_mu_size=8, scattering_mu_s_size=64, scattering_nu_size=32, order=3)


def tables(t, idx=None, orders=(2, 3, 4)):
    pick = (lambda a: a) if idx is None else (lambda a: a.reshape(-1, 4)[idx])
    out = dict(transmittance=t.transmittance.astype(np.float32), irradiance=t.irradiance.astype(np.float32),
               direct_irradiance=t.history["single"]["delta_irradiance"].astype(np.float32),
               scattering=pick(t.scattering).astype(np.float16), delta_rayleigh=pick(t.delta_rayleigh).astype(np.float16),
               delta_mie=pick(t.delta_mie).astype(np.float16), scattering_single=pick(t.history["single"]["scattering"]).astype(np.float16))
    for order in orders:
        h = t.history[order]
        out[f"o{order}_scattering_density"] = pick(h["scattering_density"]).astype(np.float16)
        out[f"o{order}_delta_multiple_scattering"] = pick(h["delta_multiple_scattering"]).astype(np.float16)
        out[f"o{order}_scattering"] = pick(h["scattering"]).astype(np.float16)
        out[f"o{order}_delta_irradiance"] = h["delta_irradiance"].astype(np.float32)
        out[f"o{order}_irradiance"] = h["irradiance"].astype(np.float32)
    return out


def more_dims(t0):
    po = O.Params(**ODD_DIMS)
    np.savez_compressed(os.path.join(HERE, "reference_odd_f32.npz"), **tables(R.precompute(po, keep_history=True)))
    print("odd dims done", time.time() - t0)
    pw = O.Params(**WIDE_DIMS)
    w = R.precompute(pw, keep_history=True)
    idx = np.sort(np.random.default_rng(4321).choice(int(np.prod(pw.s_shape[:3])), 4096, replace=False)).astype(np.int64)
    g = tables(w, idx, orders=(2, 3))
    g["idx"] = idx
    np.savez_compressed(os.path.join(HERE, "reference_wide_f32.npz"), **g)
    print("wide dims done", time.time() - t0)


def main():
    assert R.build(), "the reference checkout (/root/reference/shaders) is needed to generate these fixtures"
    t0 = time.time()
    if sys.argv[1:] == ["extra"]:               # only the two fixtures added later
        return more_dims(t0)
    ps = O.Params(**SMOKE_DIMS)
    s = R.precompute(ps, keep_history=True)
    out = tables(s)
    W, H = 48, 27
    draws, extra = synthetic.camera_sweep(24, W, H)
    for k in VIEWS_SMALL:
        depth = synthetic.analytic_depth(extra[k][0], extra[k][1], W, H)
        c, tr = R.render(ps, s.transmittance, s.scattering, O.pack_draw(draws[k].inverse_viewproj, draws[k].camera_position, draws[k].sun_direction), depth)
        out[f"view{k}_color"], out[f"view{k}_transmittance"] = c.astype(np.float32), tr.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_smoke_f32.npz"), **out)
    print("smoke dims done", time.time() - t0)

    p = O.Params()
    d = R.precompute(p, keep_history=True)
    print("default dims done", time.time() - t0)
    n_tex = int(np.prod(p.s_shape[:3]))
    idx = np.sort(np.random.default_rng(1234).choice(n_tex, 4096, replace=False)).astype(np.int64)
    g = tables(d, idx)
    g["idx"] = idx
    np.savez_compressed(os.path.join(HERE, "reference_default_f32.npz"), **g)
    print("wrote fixtures", time.time() - t0)
    more_dims(t0)


if __name__ == "__main__":
    main()
