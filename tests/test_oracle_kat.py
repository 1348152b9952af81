"""Pins the CPU oracle with the closed-form identities of SURVEY.md §8c (the reference has no golden
vectors: its only test, tests/smoke.rs, asserts "no Vulkan validation error")."""
import math

import numpy as np
import pytest

from oracle import oracle as O

from .conftest import SMOKE_DIMS

P = O.Params()
MODES = (O.F32, O.F64Q, O.F64)


def test_params_block_layout():
    b = P.pack()
    assert len(b) == 320 == O.lib().fbo_params_size()
    f = np.frombuffer(b, dtype=np.float32)
    i = np.frombuffer(b, dtype=np.int32)
    assert f[0] == np.float32(1.474) and f[3] == np.float32(0.004675)          # offsets 0, 12
    assert f[7] == 6360.0 and f[11] == 6420.0                                   # 28, 44
    assert f[15] == np.float32(0.8) and f[19] == np.float32(-0.207912)          # 60, 76
    assert list(i[23:31]) == [256, 64, 32, 128, 32, 8, 64, 16]                  # 92..120: T_mu,T_r,S_r,S_mu,S_mus,S_nu,E_mus,E_r
    assert f[32 + 8 + 1] == 1.0 and f[32 + 8 + 2] == np.float32(-0.125)         # rayleigh layer 1 at 128+32
    assert f[64] == 25.0 and f[64 + 3] == np.float32(0.066667)                  # absorption layer 0 at 256


def test_half_rounding_matches_ieee():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-70000, 70000, 2000), rng.uniform(-1, 1, 2000) * 1e-4, rng.uniform(-1, 1, 2000) * 1e-7,
                        [0.0, 65504.0, 65519.9, 65520.0, 2.0 ** -24, 2.0 ** -25, 1.5 * 2.0 ** -24, 1.0 + 2.0 ** -11]])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).astype(np.float64)
    assert np.array_equal(O.round_to_half(x), want)


@pytest.mark.parametrize("mode", MODES)
def test_transmittance_kats(mode):
    T = O.transmittance(P, mode)
    # (iii) r = top, mu = 1: zero path length
    assert np.array_equal(T[63, 0], [1, 1, 1, 1])
    # (i) r = bottom, mu = 1: vertical path, optical lengths in closed form
    def col(scale_h):   # integral of exp(-h/H) over 0..60 km
        return scale_h * (1 - math.exp(-60.0 / scale_h))
    ozone = 0.5 * 15.0 * 1.0 * 2            # triangle: 0 at 10 km, 1 at 25 km, 0 at 40 km
    tau = (np.array(P.rayleigh_scattering) * col(8.0) + np.array(P.mie_extinction) * col(1.0 / 0.833333)
           + np.array(P.absorbtion_extinction) * ozone)
    np.testing.assert_allclose(T[0, 0, :3], np.exp(-tau), rtol=2e-5)
    # (ii) horizontal ray at the ground (501-point trapezoid vs the continuous integral)
    np.testing.assert_allclose(T[0, 255, :3], [0.106448, 9.5837e-3, 5.2123e-5], rtol=2e-4)
    assert np.all(T[..., 3] == 1.0)
    assert np.all((T[..., :3] > 0) & (T[..., :3] <= 1))
    # transmittance decreases towards the horizon at every altitude
    assert np.all(np.diff(T[:, :, 0], axis=1) <= 1e-6)


@pytest.mark.parametrize("mode", MODES)
def test_direct_irradiance_kats(mode):
    T = O.transmittance(P, mode)
    dE = O.direct_irradiance(P, mode, T)
    np.testing.assert_allclose(dE[15, 63, :3], P.solar_irradiance, rtol=1e-6)   # (iv) r = top, mu_s = 1
    mu_s = 2 * np.arange(64) / 63 - 1
    assert np.all(dE[:, mu_s < -P.sun_angular_radius] == 0)
    assert np.all(dE[..., 3] == 0)


def test_constants():
    H = math.sqrt(6420.0 ** 2 - 6360.0 ** 2)
    assert abs(H - 875.6712) < 1e-3                                             # (v)
    assert abs(3 / (16 * math.pi) * 2 - 0.1193662) < 1e-7                       # (vi) P_R(1)
    g = 0.8
    pm = 3 / (8 * math.pi) * (1 - g * g) / (2 + g * g) * 2 / (1 + g * g - 2 * g) ** 1.5
    assert abs(pm - 4.0693025) < 1e-6


@pytest.mark.parametrize("mode", (O.F32, O.F64))
def test_precompute_structure(mode):
    p = O.Params(**SMOKE_DIMS)
    t4 = O.precompute(p, mode, keep_history=True)
    # (vii) alpha channels
    assert np.all(t4.delta_rayleigh[..., 3] == 0) and np.all(t4.delta_mie[..., 3] == 0)
    assert np.all(t4.delta_multiple_scattering[..., 3] == 0) and np.all(t4.scattering_density[..., 3] == 0)
    single = t4.history["single"]["scattering"]
    assert np.array_equal(t4.scattering[..., 3], single[..., 3])                # single-Mie red, untouched by later orders
    assert np.array_equal(single[..., 3], t4.delta_mie[..., 0])
    # (ix) every order adds a non-negative term
    prev = single
    for order in (2, 3, 4):
        cur = t4.history[order]["scattering"]
        assert np.all(cur[..., :3] >= prev[..., :3])
        prev = cur
    # (viii) order 1: single scattering only, irradiance identically zero
    t1 = O.precompute(O.Params(order=1, **SMOKE_DIMS), mode)
    assert np.all(t1.irradiance == 0)
    assert np.array_equal(t1.scattering, single)
    assert np.all(np.isfinite(t4.scattering)) and np.all(np.isfinite(t4.irradiance))


def test_modes_agree_where_well_conditioned():
    """fp32-as-written vs fp64: the bulk agrees to an fp16 ulp; the horizon-grazing rows do not
    (catastrophic cancellation in the reference's own formulas) — documented in DESIGN.md."""
    p = O.Params(**SMOKE_DIMS)
    a, b = O.precompute(p, O.F32), O.precompute(p, O.F64Q)
    rel = np.abs(a.scattering - b.scattering) / np.maximum(np.abs(b.scattering), 2.0 ** -14)
    assert np.median(rel) < 1e-3
    assert np.quantile(rel, 0.9) < 5e-3
    relT = np.abs(a.transmittance - b.transmittance) / np.abs(b.transmittance)
    assert np.quantile(relT, 0.99) < 1e-3


def test_sampler_is_vulkan_linear_clamp():
    """A 2x1 RGBA table sampled through GetIrradiance's path: exact texel-centre hits and edge clamping."""
    p = O.Params(irradiance_mu_s_size=2, irradiance_r_size=2)
    E = np.zeros(p.e_shape)
    E[:, 0, :3] = 1.0
    E[:, 1, :3] = 3.0
    T = O.transmittance(p, O.F64)
    pts = np.array([[0, 0, p.bottom_radius]] * 3, dtype=np.float64)
    nrm = np.array([[0, 0, 1.0]] * 3)
    sun = np.array([[math.sqrt(1 - m * m), 0, m] for m in (-1.0, 0.0, 1.0)])
    _, sky = O.sun_sky_irradiance(p, O.F64, T, E, pts, nrm, sun)
    # x_mu_s = 0, .5, 1 -> u = .25, .5, .75 -> texel centre 0, midpoint, texel centre 1 ; (1 + n.p/r)/2 = 1
    np.testing.assert_allclose(sky[:, 0], [1.0, 2.0, 3.0], rtol=1e-12)
