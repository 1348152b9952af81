import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _gpu_present() -> bool:
    """True when fb_builder_create finds an sm_100 device (the library has no CPU fallback: FB_ERR_NO_DEVICE otherwise)."""
    try:
        if not os.path.exists(os.path.join(ROOT, "fuzzyblue_b200", "csrc", "libfuzzyblue_b200.so")):
            import __graft_entry__ as g
            g.build()
        from fuzzyblue_b200 import api
        b = api.Builder(0)
        b.close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a B200: on a machine without one they are skipped (not failed with FB_ERR_NO_DEVICE), so a
    plain `pytest tests` is green on the CPU box; the driver runs them with `-m gpu` on the GPU box."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _gpu_present():
        return
    skip = pytest.mark.skip(reason="no sm_100 device visible (FB_ERR_NO_DEVICE)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Make sure both shared libraries exist (cross-compiles here; prebuilt on the GPU box)."""
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "fuzzyblue_b200", "csrc", "libfuzzyblue_b200.so")
    orc = os.path.join(ROOT, "oracle", "libfb_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        g.build()


SMOKE_DIMS = dict(scattering_r_size=8, scattering_mu_size=32, scattering_mu_s_size=8, scattering_nu_size=2)    # tests/smoke.rs:136-142
DUMP_DIMS = dict(scattering_r_size=16, scattering_mu_size=64, scattering_mu_s_size=16, scattering_nu_size=4)   # examples/dump.rs:101-107


# neither powers of two nor multiples of a warp anywhere (tests/golden/reference_odd_f32.npz holds the reference's output)
ODD_DIMS = dict(scattering_r_size=7, scattering_mu_size=22, scattering_mu_s_size=11, scattering_nu_size=3,
                transmittance_mu_size=100, transmittance_r_size=33, irradiance_mu_s_size=20, irradiance_r_size=9)


# a thin cut of the high-resolution config (BASELINE.json configs[2]: nu 32, mu_s 128): rows wider than one CTA
WIDE_DIMS = dict(scattering_r_size=4, scattering_mu_size=8, scattering_mu_s_size=64, scattering_nu_size=32, order=3)


# more altitude levels than one density launch carries ground normals for (64), and an irradiance table that is not 64 wide
TALL_DIMS = dict(scattering_r_size=72, scattering_mu_size=4, scattering_mu_s_size=4, scattering_nu_size=2,
                 irradiance_mu_s_size=40, irradiance_r_size=8, order=3)


@pytest.fixture(scope="session")
def oracle_tall_f32():
    from oracle import oracle as O
    return O.precompute(O.Params(**TALL_DIMS), O.F32, keep_history=True)


@pytest.fixture(scope="session")
def oracle_wide_f32():
    from oracle import oracle as O
    t = O.precompute(O.Params(**WIDE_DIMS), O.F32, keep_history=True)
    assert_equals_reference_golden(t, "reference_wide_f32.npz")
    return t


@pytest.fixture(scope="session")
def oracle_smoke_f32():
    from oracle import oracle as O
    t = O.precompute(O.Params(**SMOKE_DIMS), O.F32, keep_history=True)
    assert_equals_reference_golden(t, "reference_smoke_f32.npz")
    return t


@pytest.fixture(scope="session")
def oracle_dump_f32():
    from oracle import oracle as O
    return O.precompute(O.Params(**DUMP_DIMS), O.F32, keep_history=True)


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def assert_equals_reference_golden(tables, fixture: str):
    """`tables` (an oracle run with history) must reproduce, bit for bit, the golden vectors the REFERENCE'S OWN SHADERS
    produced (tests/golden/make_reference_golden.py): whole tables where the fixture holds them, the seeded texels
    otherwise.  Every expectation the GPU tests derive from the oracle is thereby the reference's own output."""
    import numpy as np
    g = np.load(os.path.join(GOLDEN_DIR, fixture))
    idx = g["idx"] if "idx" in g.files else None
    pick = (lambda a: a) if idx is None else (lambda a: a.reshape(-1, 4)[idx])
    assert np.array_equal(tables.transmittance.astype(np.float32), g["transmittance"])
    assert np.array_equal(tables.irradiance.astype(np.float32), g["irradiance"])
    assert np.array_equal(pick(tables.scattering).astype(np.float16), g["scattering"])
    assert np.array_equal(pick(tables.delta_rayleigh).astype(np.float16), g["delta_rayleigh"])
    assert np.array_equal(pick(tables.delta_mie).astype(np.float16), g["delta_mie"])
    for order in sorted(int(k[1]) for k in g.files if k.endswith("_scattering_density")):
        h = tables.history[order]
        for k in ("scattering_density", "delta_multiple_scattering", "scattering"):
            assert np.array_equal(pick(h[k]).astype(np.float16), g[f"o{order}_{k}"]), (fixture, order, k)
        for k in ("delta_irradiance", "irradiance"):
            assert np.array_equal(h[k].astype(np.float32), g[f"o{order}_{k}"]), (fixture, order, k)


@pytest.fixture(scope="session")
def oracle_default_f32():
    """BASELINE.json configs[1] in full: the fp32 oracle's 4-order precompute at the default dims with every intermediate
    image of every order (15-40 s on the GPU box's host cores; used by the -m gpu parity tests only).  Checked against
    the reference's own golden vectors before any kernel is compared with it."""
    from oracle import oracle as O
    t = O.precompute(O.Params(), O.F32, keep_history=True)
    assert_equals_reference_golden(t, "reference_default_f32.npz")
    return t
