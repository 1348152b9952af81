"""Cross-check of the C++ oracle (fp64, no quantisation) against an independent scalar numpy restatement of the shaders
(oracle/numpy_check.py) on a handful of texels of every stage.  Agreement to ~1e-11 means neither transcription has a
typo the other lacks; it does not pin either against the reference (which has no golden vectors)."""
import numpy as np
import pytest

from oracle import numpy_check as N
from oracle import oracle as O

DIMS = dict(scattering_r_size=4, scattering_mu_size=8, scattering_mu_s_size=4, scattering_nu_size=3,
            transmittance_mu_size=32, transmittance_r_size=8, irradiance_mu_s_size=8, irradiance_r_size=4)
P = O.Params(order=3, **DIMS)
A = N.Atm(P)
RTOL = 1e-9


@pytest.fixture(scope="module")
def tables():
    return O.precompute(P, O.F64, keep_history=True)


def texels():
    W = P.scattering_nu_size * P.scattering_mu_s_size
    return [(0, 0, 0), (5, 3, 1), (W - 1, P.scattering_mu_size - 1, P.scattering_r_size - 1), (7, 4, 2), (2, 6, 3), (9, 1, 2)]


def test_transmittance(tables):
    for x, y in ((0, 0), (31, 0), (0, 7), (13, 4), (30, 6)):
        np.testing.assert_allclose(N.transmittance_texel(A, x, y), tables.transmittance[y, x, :3], rtol=RTOL)


def test_single_scattering(tables):
    for x, y, z in texels():
        ray, mie = N.single_scattering_texel(A, tables.transmittance, x, y, z)
        np.testing.assert_allclose(ray, tables.delta_rayleigh[z, y, x, :3], rtol=RTOL, atol=1e-300)
        np.testing.assert_allclose(mie, tables.delta_mie[z, y, x, :3], rtol=RTOL, atol=1e-300)


@pytest.mark.parametrize("order", [2, 3])
def test_scattering_density(tables, order):
    prev = tables.history["single"] if order == 2 else tables.history[order - 1]
    dMS = prev.get("delta_multiple_scattering", np.zeros(P.s_shape))
    for x, y, z in texels()[:4]:
        got = N.scattering_density_texel(A, tables.transmittance, tables.delta_rayleigh, tables.delta_mie, dMS,
                                         prev["delta_irradiance"], x, y, z, order)
        np.testing.assert_allclose(got, tables.history[order]["scattering_density"][z, y, x, :3], rtol=RTOL, atol=1e-300)


@pytest.mark.parametrize("order", [2, 3])
def test_indirect_irradiance(tables, order):
    prev = tables.history["single"] if order == 2 else tables.history[order - 1]
    dMS = prev.get("delta_multiple_scattering", np.zeros(P.s_shape))
    for x, y in ((0, 0), (7, 3), (4, 1), (6, 2)):
        got = N.indirect_irradiance_texel(A, tables.delta_rayleigh, tables.delta_mie, dMS, x, y, order - 1)
        np.testing.assert_allclose(got, tables.history[order]["delta_irradiance"][y, x, :3], rtol=RTOL, atol=1e-300)


def test_multiple_scattering(tables):
    h = tables.history[2]
    for x, y, z in texels():
        got, nu = N.multiple_scattering_texel(A, tables.transmittance, h["scattering_density"], x, y, z)
        np.testing.assert_allclose(got, h["delta_multiple_scattering"][z, y, x, :3], rtol=RTOL, atol=1e-300)
        want = tables.history["single"]["scattering"][z, y, x, :3] + got / N.rayleigh_phase(nu)
        np.testing.assert_allclose(want, h["scattering"][z, y, x, :3], rtol=RTOL, atol=1e-300)


def test_render_pixels(tables):
    """render_sky.frag on a few geometry pixels of three views (ground below, at altitude, from space)."""
    from fuzzyblue_b200 import synthetic
    W, H = 32, 18
    draws, extra = synthetic.camera_sweep(14, W, H)
    checked = 0
    for k in (1, 8, 13):
        depth = synthetic.analytic_depth(extra[k][0], extra[k][1], W, H)
        pd = O.pack_draw(draws[k].inverse_viewproj, draws[k].camera_position, draws[k].sun_direction)
        oc, ot = O.render(P, O.F64, tables.transmittance, tables.scattering, pd, depth)
        ys, xs = np.nonzero(depth > 0)
        for j in range(0, len(ys), max(1, len(ys) // 6)):
            c, t = N.render_pixel(A, tables.transmittance, tables.scattering, pd, depth[ys[j], xs[j]], xs[j], ys[j], W, H)
            # the colour is a difference of nearly equal look-ups: compare at the scale of its operands
            np.testing.assert_allclose(c, oc[ys[j], xs[j], :3], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(t, ot[ys[j], xs[j], :3], rtol=1e-7)
            checked += 1
    assert checked >= 12
