"""Known-answer tests for the 3-D stages (single scattering, scattering density, indirect irradiance, multiple
scattering): closed forms of tests/analytic.py asserted on the CPU oracle (all arithmetic modes) and, through
fb_pending_upload + fb_pending_run_stage, on the CUDA kernels of both families.

The reference ships no golden vectors (tests/smoke.rs:169-171 asserts only "no validation error"), so these
identities are the pin that does not come from the oracle itself: each feeds a stage constant input tables, for which
the stage's quadrature collapses to a formula in the texel geometry (shaders/scattering_density.comp:39-106,
indirect_irradiance.comp:27-44, multiple_scattering.comp:22-52, single_scattering.comp:10-70).

Tolerances.  fp16 storage: |a - b| <= 1e-3 max(|b|, 2^-14); fp32 storage: 1e-3 relative.  In fp64 arithmetic (oracle
modes 1, 2) every texel must hold it.  In fp32 arithmetic (oracle mode 0 = the shaders as written, and the CUDA kernels)
the density and irradiance identities hold on every texel too; the two ray-marching stages re-derive the ray length
from (r, mu) through `-r mu -+ sqrt(r^2 (mu^2 - 1) + R^2)`, a difference of 4e7-sized numbers whose fp32 rounding moves
horizon-grazing rays by up to a few per cent: there the gate is >= 97 % of texels within 1e-3 and
none beyond 8e-2, and the CUDA result must match the fp32 oracle on the same inputs within 1e-3 on EVERY texel.
"""
import math

import numpy as np
import pytest

from oracle import oracle as O

from . import analytic as AN
from .conftest import DUMP_DIMS

F16_FLOOR = 2.0 ** -14
P = O.Params(**DUMP_DIMS)

C_RGB = np.array([64.0, 48.0, 80.0])        # uniform incident radiance (fp16-exact; keeps the outputs in the normal fp16 range)
E_RGB = np.array([512.0, 768.0, 640.0])     # uniform ground irradiance
J_RGB = np.array([1 / 64.0, 1 / 32.0, 3 / 128.0])   # uniform scattering density
E0_RGB = np.array([0.25, 0.5, 0.125])       # irradiance already accumulated


def err16(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), F16_FLOOR)


def err32(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


def strict(name, e, tol=1e-3):
    print(f"{name}: max {e.max():.3e}")
    assert e.max() <= tol, f"{name}: max {e.max():.3e} at {np.unravel_index(int(e.argmax()), e.shape)}"


def grazing(name, e):
    frac = float((e > 1e-3).mean())
    print(f"{name}: max {e.max():.3e}, beyond 1e-3: {frac:.2e}")
    assert frac <= 3e-2 and e.max() <= 8e-2, f"{name}: max {e.max():.3e}, fraction beyond 1e-3 {frac:.2e}"


def const_table(shape, rgb, alpha=0.0):
    t = np.zeros(shape)
    t[..., :3] = rgb
    t[..., 3] = alpha
    return t


ONES_T = const_table(P.t_shape, 1.0, 1.0)


# ------------------------------------------------------------------------------------------------------------------
# the closed forms themselves
# ------------------------------------------------------------------------------------------------------------------
def test_phase_functions_integrate_to_one_under_the_16x32_rule():
    _, sums = AN.density_uniform(P, C_RGB, E_RGB)
    assert np.abs(sums["R"] - 1.0).max() < 3e-3         # Rayleigh: smooth, the midpoint rule is within 0.3 %
    assert np.abs(sums["M"] - 1.0).max() < 0.35         # Mie g = 0.8: the forward lobe is under-resolved by 512 directions
    assert (sums["Rg"] <= sums["R"]).all() and (sums["Rg"] > 0).any()


def test_hemisphere_cosine_sum_is_pi():
    _, s = AN.indirect_irradiance_uniform(C_RGB)
    assert abs(s / math.pi - 1.0) < 2e-3 and s > math.pi        # the midpoint rule overshoots by h^2 / 6 = 1.6e-3 (h = pi / 32)


def test_trapezoid_of_the_density_profile_matches_quad():
    """Near-vertical rays with the sun overhead (visibility 1 all along the ray): the 51-node trapezoid of rho_R along
    the ray against scipy.integrate.quad of the same integrand (quadrature error of the rule itself <= 1 %)."""
    from scipy.integrate import quad
    ray, _, g = AN.single_scattering_unit_transmittance(P)
    NR, NMU, NNU, NMS = g.shape
    ray = ray.reshape(NR, NMU, NNU, NMS, 3)
    sol_beta = np.float64(np.float32(P.solar_irradiance[0])) * np.float64(np.float32(P.rayleigh_scattering[0]))
    n = 0
    for z in (0, NR // 2, NR - 2):
        for y in (NMU - 1, NMU - 2, NMU // 2 + 3, 0):
            r, mu, d = float(g.r[z, 0, 0, 0]), float(g.mu[z, y, 0, 0]), float(g.d[z, y, 0, 0])
            f = lambda s: math.exp(-0.125 * max(math.sqrt(s * s + 2 * r * mu * s + r * r) - g.bot, 0.0))
            want = sol_beta * quad(f, 0.0, d, epsabs=0, epsrel=1e-10)[0]
            got = ray[z, y, NNU - 1, NMS - 1, 0]                # mu_s = 1 (sun at the zenith), nu clamped to mu
            assert abs(got - want) <= 1e-2 * want, (z, y, got, want)
            n += 1
    assert n == 12


# ------------------------------------------------------------------------------------------------------------------
# the oracle against the closed forms (CPU)
# ------------------------------------------------------------------------------------------------------------------
MODES = [O.F32, O.F64Q, O.F64]


@pytest.mark.parametrize("mode", MODES)
def test_oracle_density_uniform_radiance(mode):
    want, _ = AN.density_uniform(P, C_RGB, E_RGB)
    zero3 = np.zeros(P.s_shape)
    got = O.scattering_density(P, mode, 3, ONES_T, zero3, zero3, const_table(P.s_shape, C_RGB), const_table(P.e_shape, E_RGB))
    strict(f"oracle density KAT mode {mode}", err16(got[..., :3], want))
    assert np.all(got[..., 3] == 0)


@pytest.mark.parametrize("mode", MODES)
def test_oracle_density_order2_uses_both_single_tables(mode):
    """order 2: radiance = rayleigh * P_R(nu1) + mie * P_M(nu1) (scattering.h:165-173).  With delta_mie == 0 and
    delta_rayleigh == c the incident radiance is c P_R(nu1) -- not constant, but its integral against the Rayleigh
    kernel is a double midpoint sum the test evaluates directly for the mu_s = 1 column (omega_s = zenith: nu1 = cos theta)."""
    g = AN.Geometry(P)
    N = 16
    th = (np.arange(N) + 0.5) * math.pi / N
    ph = (np.arange(2 * N) + 0.5) * math.pi / N
    ct, st, cp = np.cos(th)[:, None], np.sin(th)[:, None], np.cos(ph)[None, :]
    mu = g.mu[:, :, 0, 0]
    nu2 = np.sqrt(1 - mu * mu)[..., None, None] * (cp * st) + mu[..., None, None] * ct
    dw = (math.pi / N) ** 2 * st
    gM = float(np.float32(P.mie_phase_function_g))
    inc = AN.rayleigh_phase(ct)                                               # nu1 = cos(theta) when the sun is at the zenith
    h = g.r[:, :, 0, 0] - g.bot
    k = (np.float64(np.float32(P.rayleigh_scattering[0])) * AN.profile_density(P.rayleigh_density, h)[..., None, None] * AN.rayleigh_phase(nu2)
         + np.float64(np.float32(P.mie_scattering[0])) * AN.profile_density(P.mie_density, h)[..., None, None] * AN.mie_phase(gM, nu2))
    want = C_RGB[0] * (inc * k * dw).sum((-1, -2))                            # [R][MU], red channel
    zero3 = np.zeros(P.s_shape)
    got = O.scattering_density(P, mode, 2, ONES_T, const_table(P.s_shape, C_RGB), zero3, zero3, np.zeros(P.e_shape))
    NR, NMU, NNU, NMS = g.shape
    col = got.reshape(NR, NMU, NNU, NMS, 4)[:, :, :, NMS - 1, 0]              # mu_s texel NMS-1: mu_s = 1 exactly
    assert float(g.mu_s[0, 0, 0, NMS - 1]) == 1.0
    # fp32: `r - bottom` of the lowest levels carries 1e-3-relative rounding into d and mu (scattering.h:84-92), and the Mie
    # lobe turns a 2e-5 error of nu2 into 1.3e-3 (d ln P_M / d nu = 3 g / (1 + g^2 - 2 g nu) = 60 at nu = 1)
    for kk in range(NNU):
        strict(f"oracle order-2 density KAT mode {mode} nu slice {kk}", err16(col[:, :, kk], want), 3e-3 if mode == O.F32 else 1e-3)


@pytest.mark.parametrize("mode", MODES)
def test_oracle_indirect_irradiance_uniform_radiance(mode):
    want, _ = AN.indirect_irradiance_uniform(C_RGB)
    zero3 = np.zeros(P.s_shape)
    dE, E = O.indirect_irradiance(P, mode, 2, zero3, zero3, const_table(P.s_shape, C_RGB), const_table(P.e_shape, E0_RGB))
    strict(f"oracle indirect irradiance KAT mode {mode}", err32(dE[..., :3], np.broadcast_to(want, dE[..., :3].shape)), 1e-5)
    strict(f"oracle irradiance accumulation mode {mode}", err32(E[..., :3], np.broadcast_to(E0_RGB + want, E[..., :3].shape)), 1e-5)
    strict("vs pi c", err32(dE[..., :3], np.broadcast_to(math.pi * C_RGB, dE[..., :3].shape)), 2e-3)


@pytest.mark.parametrize("mode", MODES)
def test_oracle_multiple_scattering_uniform_density(mode):
    want_d, want_inc, _ = AN.multiple_scattering_uniform(P, J_RGB)
    S0 = const_table(P.s_shape, 0.0, 0.375)
    dMS, S = O.multiple_scattering(P, mode, ONES_T, const_table(P.s_shape, J_RGB), S0)
    gate = grazing if mode == O.F32 else strict
    gate(f"oracle multiple scattering KAT mode {mode}: delta", err16(dMS[..., :3], want_d))
    gate(f"oracle multiple scattering KAT mode {mode}: scattering", err16(S[..., :3], want_inc))
    assert np.all(dMS[..., 3] == 0) and np.all(S[..., 3] == 0.375)


@pytest.mark.parametrize("mode", MODES)
def test_oracle_single_scattering_unit_transmittance(mode):
    want_r, want_m, _ = AN.single_scattering_unit_transmittance(P)
    dR, dM, S = O.single_scattering(P, mode, ONES_T)
    gate = grazing if mode == O.F32 else strict
    gate(f"oracle single scattering KAT mode {mode}: rayleigh", err16(dR[..., :3], want_r))
    gate(f"oracle single scattering KAT mode {mode}: mie", err16(dM[..., :3], want_m))
    assert np.array_equal(S[..., :3], dR[..., :3]) and np.array_equal(S[..., 3], dM[..., 0])


# ------------------------------------------------------------------------------------------------------------------
# the CUDA kernels against the closed forms and against the fp32 oracle on the same inputs (GPU, through the C ABI)
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module", params=["fast", "reference"])
def pending(request):
    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import api
    b = fb.Builder(0, kernels=api.KERNELS_FAST if request.param == "fast" else api.KERNELS_REFERENCE)
    pend = fb.Atmosphere.allocate(b, fb.Parameters(**DUMP_DIMS))
    yield pend
    pend.close()
    b.close()


def _upload_all(pend, tables):
    from fuzzyblue_b200 import api
    zeros3, zeros2 = np.zeros(P.s_shape), np.zeros(P.e_shape)
    base = {api.IMAGE_TRANSMITTANCE: ONES_T, api.IMAGE_IRRADIANCE: zeros2, api.IMAGE_DELTA_IRRADIANCE: zeros2,
            api.IMAGE_SCATTERING: zeros3, api.IMAGE_DELTA_RAYLEIGH: zeros3, api.IMAGE_DELTA_MIE: zeros3,
            api.IMAGE_SCATTERING_DENSITY: zeros3, api.IMAGE_DELTA_MULTIPLE_SCATTERING: zeros3}
    base.update(tables)
    for image, host in base.items():
        pend.upload(image, host)


@pytest.mark.gpu
def test_cuda_density_uniform_radiance(pending):
    from fuzzyblue_b200 import api
    want, _ = AN.density_uniform(P, C_RGB, E_RGB)
    _upload_all(pending, {api.IMAGE_DELTA_MULTIPLE_SCATTERING: const_table(P.s_shape, C_RGB),
                          api.IMAGE_DELTA_IRRADIANCE: const_table(P.e_shape, E_RGB)})
    pending.run_stage(api.STAGE_SCATTERING_DENSITY, order=3)
    got = pending.download(api.IMAGE_SCATTERING_DENSITY)
    strict("CUDA density KAT", err16(got[..., :3], want))
    assert np.all(got[..., 3] == 0)


@pytest.mark.gpu
def test_cuda_density_order2_single_tables(pending):
    from fuzzyblue_b200 import api
    zero3 = np.zeros(P.s_shape)
    ref = O.scattering_density(P, O.F32, 2, ONES_T, const_table(P.s_shape, C_RGB), const_table(P.s_shape, C_RGB[::-1]), zero3,
                               const_table(P.e_shape, E_RGB))
    _upload_all(pending, {api.IMAGE_DELTA_RAYLEIGH: const_table(P.s_shape, C_RGB), api.IMAGE_DELTA_MIE: const_table(P.s_shape, C_RGB[::-1]),
                          api.IMAGE_DELTA_IRRADIANCE: const_table(P.e_shape, E_RGB)})
    pending.run_stage(api.STAGE_SCATTERING_DENSITY, order=2)
    strict("CUDA order-2 density vs oracle (pinned by its own order-2 KAT)", err16(pending.download(api.IMAGE_SCATTERING_DENSITY), ref))


@pytest.mark.gpu
def test_cuda_indirect_irradiance_uniform_radiance(pending):
    from fuzzyblue_b200 import api
    want, _ = AN.indirect_irradiance_uniform(C_RGB)
    _upload_all(pending, {api.IMAGE_DELTA_MULTIPLE_SCATTERING: const_table(P.s_shape, C_RGB), api.IMAGE_IRRADIANCE: const_table(P.e_shape, E0_RGB)})
    pending.run_stage(api.STAGE_INDIRECT_IRRADIANCE, order=2)
    dE, E = pending.download(api.IMAGE_DELTA_IRRADIANCE), pending.download(api.IMAGE_IRRADIANCE)
    strict("CUDA indirect irradiance KAT", err32(dE[..., :3], np.broadcast_to(want, dE[..., :3].shape)), 1e-5)
    strict("CUDA irradiance accumulation", err32(E[..., :3], np.broadcast_to(E0_RGB + want, E[..., :3].shape)), 1e-5)


@pytest.mark.gpu
def test_cuda_multiple_scattering_uniform_density(pending):
    from fuzzyblue_b200 import api
    want_d, want_inc, _ = AN.multiple_scattering_uniform(P, J_RGB)
    S0 = const_table(P.s_shape, 0.0, 0.375)
    _upload_all(pending, {api.IMAGE_SCATTERING_DENSITY: const_table(P.s_shape, J_RGB), api.IMAGE_SCATTERING: S0})
    pending.run_stage(api.STAGE_MULTIPLE_SCATTERING)
    dMS, S = pending.download(api.IMAGE_DELTA_MULTIPLE_SCATTERING), pending.download(api.IMAGE_SCATTERING)
    grazing("CUDA multiple scattering KAT: delta", err16(dMS[..., :3], want_d))
    grazing("CUDA multiple scattering KAT: scattering", err16(S[..., :3], want_inc))
    odMS, oS = O.multiple_scattering(P, O.F32, ONES_T, const_table(P.s_shape, J_RGB), S0)
    strict("CUDA multiple scattering vs fp32 oracle, same inputs: delta", err16(dMS, odMS))
    strict("CUDA multiple scattering vs fp32 oracle, same inputs: scattering", err16(S, oS))


@pytest.mark.gpu
def test_cuda_single_scattering_unit_transmittance(pending):
    from fuzzyblue_b200 import api
    want_r, want_m, _ = AN.single_scattering_unit_transmittance(P)
    _upload_all(pending, {})
    pending.run_stage(api.STAGE_SINGLE_SCATTERING)
    dR, dM = pending.download(api.IMAGE_DELTA_RAYLEIGH), pending.download(api.IMAGE_DELTA_MIE)
    grazing("CUDA single scattering KAT: rayleigh", err16(dR[..., :3], want_r))
    grazing("CUDA single scattering KAT: mie", err16(dM[..., :3], want_m))
    odR, odM, _ = O.single_scattering(P, O.F32, ONES_T)
    strict("CUDA single scattering vs fp32 oracle, same inputs: rayleigh", err16(dR, odR))
    strict("CUDA single scattering vs fp32 oracle, same inputs: mie", err16(dM, odM))
