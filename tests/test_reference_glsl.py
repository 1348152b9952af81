"""The oracle against THE REFERENCE ITSELF: /root/reference/shaders/*.h, *.comp and render_sky.frag compiled as C++
(oracle/glsl_ref: a GLSL-subset header + a purely syntactic translator, the shader logic is read from the reference
checkout at build time) and executed on the CPU.  Every stage, the whole per-order schedule and the sky evaluation must
agree with oracle/fb_oracle.cpp (fp32 mode) BIT FOR BIT on identical inputs.

Runs wherever oracle/_ref/libfb_glsl_ref.so exists: it is built in this container by __graft_entry__.build() (where
/root/reference is present) and travels to the GPU box with the snapshot; tests/test_golden_cpu.py holds the same pin
as committed vectors for machines that have neither."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_glsl as R

from .conftest import DUMP_DIMS, SMOKE_DIMS

pytestmark = pytest.mark.skipif(not (R.build() and R.available()), reason="oracle/_ref/libfb_glsl_ref.so not built (needs /root/reference)")


def same(a, b, what):
    assert a.shape == b.shape, what
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{what}: {int(bad.sum())} of {a.size} values differ, max abs {np.nanmax(np.abs(a - b)):.3e}"


def history_same(a, b, orders):
    same(a.transmittance, b.transmittance, "transmittance")
    same(a.history["single"]["delta_irradiance"], b.history["single"]["delta_irradiance"], "direct irradiance")
    same(a.delta_rayleigh, b.delta_rayleigh, "delta_rayleigh")
    same(a.delta_mie, b.delta_mie, "delta_mie")
    same(a.history["single"]["scattering"], b.history["single"]["scattering"], "scattering (single)")
    for o in orders:
        for k in ("scattering_density", "delta_irradiance", "irradiance", "delta_multiple_scattering", "scattering"):
            same(a.history[o][k], b.history[o][k], f"order {o} {k}")
    same(a.irradiance, b.irradiance, "irradiance")
    same(a.scattering, b.scattering, "scattering")


ODD = dict(scattering_r_size=3, scattering_mu_size=6, scattering_mu_s_size=5, scattering_nu_size=4, order=3, transmittance_mu_size=37,
           transmittance_r_size=11, irradiance_mu_s_size=13, irradiance_r_size=5)


@pytest.mark.parametrize("dims,orders", [(SMOKE_DIMS, (2, 3, 4)), (DUMP_DIMS, (2, 3, 4)), (ODD, (2, 3))], ids=["smoke", "dump", "odd"])
def test_whole_precompute_bit_for_bit(dims, orders):
    p = O.Params(**dims)
    history_same(O.precompute(p, O.F32, keep_history=True), R.precompute(p, keep_history=True), orders)


@pytest.mark.parametrize("index", [0, 1, 2])
def test_randomised_atmospheres_bit_for_bit(index):
    """BASELINE.json config 4 physics (Earth-to-Mars radii, random Rayleigh / Mie / ozone / albedo / sun size)."""
    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import synthetic
    q = synthetic.random_atmospheres(3, base=fb.Parameters(order=3, **SMOKE_DIMS), seed=99)[index]
    p = O.Params(order=3, **SMOKE_DIMS)
    for name in ("solar_irradiance", "sun_angular_radius", "bottom_radius", "top_radius", "rayleigh_scattering", "mie_scattering",
                 "mie_extinction", "mie_phase_function_g", "absorbtion_extinction", "ground_albedo", "mu_s_min"):
        setattr(p, name, getattr(q, name))
    for name in ("rayleigh_density", "mie_density", "absorbtion_density"):
        setattr(p, name, tuple(O.Layer(l.width, l.exp_term, l.exp_scale, l.linear_term, l.constant_term) for l in getattr(q, name).layers))
    history_same(O.precompute(p, O.F32, keep_history=True), R.precompute(p, keep_history=True), (2, 3))


def test_sky_evaluation_bit_for_bit_over_the_sweep():
    """render_sky.frag over 24 synthetic views (ground, sky, cameras in space): every pixel of both outputs, including
    the NaN transmittance the shader itself produces for downward sky rays (inf - inf, transmittance.h:43)."""
    from fuzzyblue_b200 import synthetic
    p = O.Params(**DUMP_DIMS)
    t = O.precompute(p, O.F32)
    W, H = 96, 54
    draws, extra = synthetic.camera_sweep(24, W, H)
    n_nan = n_ground = 0
    for k in range(24):
        depth = synthetic.analytic_depth(extra[k][0], extra[k][1], W, H)
        d = O.pack_draw(draws[k].inverse_viewproj, draws[k].camera_position, draws[k].sun_direction)
        oc, ot = O.render(p, O.F32, t.transmittance, t.scattering, d, depth)
        rc, rt = R.render(p, t.transmittance, t.scattering, d, depth)
        same(oc, rc, f"view {k} colour")
        same(ot, rt, f"view {k} transmittance")
        n_nan += int(np.isnan(rt).sum())
        n_ground += int((depth > 0).sum())
    assert n_nan > 0 and n_ground > 1000


def test_sky_evaluation_bit_for_bit_with_three_nu_knots():
    """The same with a nu axis that is not a power of two (x / nu_size is then a true division in scattering.h:149-151) and
    odd sizes everywhere: 8 views of the sweep on the odd-dims tables."""
    from fuzzyblue_b200 import synthetic
    from .conftest import ODD_DIMS
    p = O.Params(**ODD_DIMS)
    t = O.precompute(p, O.F32)
    W, H = 64, 36
    draws, extra = synthetic.camera_sweep(24, W, H)
    for k in (0, 3, 5, 9, 11, 13, 17, 22):
        depth = synthetic.analytic_depth(extra[k][0], extra[k][1], W, H)
        d = O.pack_draw(draws[k].inverse_viewproj, draws[k].camera_position, draws[k].sun_direction)
        oc, ot = O.render(p, O.F32, t.transmittance, t.scattering, d, depth)
        rc, rt = R.render(p, t.transmittance, t.scattering, d, depth)
        same(oc, rc, f"view {k} colour")
        same(ot, rt, f"view {k} transmittance")


def test_default_dims_stages_on_sampled_texels_bit_for_bit():
    """BASELINE.json configs[1] dims: each 3-D stage on 1024 seeded texels of the full-size tables (identical inputs)."""
    p = O.Params()
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(int(np.prod(p.s_shape[:3])), 1024, replace=False)).astype(np.int64)
    T = O.transmittance(p, O.F32)
    same(T, R.transmittance(p), "transmittance")
    dE = O.direct_irradiance(p, O.F32, T)
    same(dE, R.direct_irradiance(p, T), "direct irradiance")
    for a, b, name in zip(O.single_scattering(p, O.F32, T, idx), R.single_scattering(p, T, idx), ("delta_rayleigh", "delta_mie", "scattering")):
        same(a, b, name)
    # smooth synthetic full-size inputs for the later stages (fp16-exact values): both implementations read the same tables
    z, y, x = np.meshgrid(np.arange(p.s_shape[0]), np.arange(p.s_shape[1]), np.arange(p.s_shape[2]), indexing="ij")
    tab = np.zeros(p.s_shape)
    for c in range(3):
        tab[..., c] = np.float16(1e-2 * (1.0 + 0.3 * np.sin(0.11 * x + c) + 0.2 * np.cos(0.07 * y) + 0.1 * z / p.s_shape[0])).astype(np.float64)
    for order in (2, 3):
        same(O.scattering_density(p, O.F32, order, T, tab, 0.5 * tab, 2.0 * tab, dE, idx),
             R.scattering_density(p, order, T, tab, 0.5 * tab, 2.0 * tab, dE, idx), f"scattering_density order {order}")
    for a, b, name in zip(O.multiple_scattering(p, O.F32, T, tab, tab, idx), R.multiple_scattering(p, T, tab, tab, idx),
                          ("delta_multiple_scattering", "scattering")):
        same(a, b, name)
    for order in (1, 2):
        for a, b, name in zip(O.indirect_irradiance(p, O.F32, order, tab, 0.5 * tab, 2.0 * tab, dE),
                              R.indirect_irradiance(p, order, tab, 0.5 * tab, 2.0 * tab, dE), ("delta_irradiance", "irradiance")):
            same(a, b, f"{name} order {order}")


def test_committed_golden_vectors_are_what_the_reference_produces():
    """tests/golden/reference_smoke_f32.npz must be reproducible from the reference checkout."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_smoke_f32.npz"))
    t = R.precompute(O.Params(**SMOKE_DIMS), keep_history=True)
    assert np.array_equal(t.scattering.astype(np.float16), g["scattering"])
    assert np.array_equal(t.irradiance.astype(np.float32), g["irradiance"])
    assert np.array_equal(t.history[3]["scattering_density"].astype(np.float16), g["o3_scattering_density"])
    from .conftest import ODD_DIMS, WIDE_DIMS
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_odd_f32.npz"))
    t = R.precompute(O.Params(**ODD_DIMS), keep_history=True)
    assert np.array_equal(t.scattering.astype(np.float16), g["scattering"]) and np.array_equal(t.irradiance.astype(np.float32), g["irradiance"])
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_wide_f32.npz"))
    t = R.precompute(O.Params(**WIDE_DIMS), keep_history=True)
    assert np.array_equal(t.scattering.reshape(-1, 4)[g["idx"]].astype(np.float16), g["scattering"])
