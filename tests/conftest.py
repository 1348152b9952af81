import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


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
    return O.precompute(O.Params(**WIDE_DIMS), O.F32, keep_history=True)


@pytest.fixture(scope="session")
def oracle_smoke_f32():
    from oracle import oracle as O
    return O.precompute(O.Params(**SMOKE_DIMS), O.F32, keep_history=True)


@pytest.fixture(scope="session")
def oracle_dump_f32():
    from oracle import oracle as O
    return O.precompute(O.Params(**DUMP_DIMS), O.F32, keep_history=True)
