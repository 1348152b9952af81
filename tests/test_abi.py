"""The C-ABI library loads and exports exactly what include/fuzzyblue.h declares (no compute here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import fuzzyblue_b200 as fb
from fuzzyblue_b200 import api
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fuzzyblue.h")).read()
    return sorted(set(re.findall(r"FB_API[^;(]*?\b(fb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 40
    L = ctypes.CDLL(api.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in fuzzyblue.h but not exported"
    assert sorted(api.ABI) == syms, "python binding table and header disagree"


def test_struct_layouts():
    assert ctypes.sizeof(api.FbParams) == 320 and ctypes.sizeof(api.FbDrawParams) == 92
    assert api.FbParams.transmittance_mu_size.offset == 92
    assert api.FbParams.irradiance_r_size.offset == 120
    assert api.FbParams.rayleigh_density.offset == 128
    assert api.FbParams.mie_density.offset == 192
    assert api.FbParams.absorption_density.offset == 256
    assert api.FbDrawParams.camera_position.offset == 64 and api.FbDrawParams.sun_direction.offset == 80


def test_default_parameters_agree_everywhere():
    raw, order = fb.Parameters.default_raw()
    assert order == 4 == fb.Parameters().order
    assert bytes(raw) == bytes(fb.Parameters().raw()) == O.Params().pack()


def test_extents():
    p = fb.Parameters()
    assert p.transmittance_extent() == (256, 64) and p.irradiance_extent() == (64, 16)
    assert p.scattering_extent() == (256, 128, 32)
    e = api.FbExtent3D()
    raw = p.raw()
    assert api._lib().fb_params_scattering_extent(ctypes.byref(raw), ctypes.byref(e)) == 0
    assert (e.width, e.height, e.depth) == (256, 128, 32)


@pytest.mark.parametrize("bad", [dict(scattering_nu_size=1), dict(scattering_mu_size=33), dict(transmittance_r_size=0),
                                 dict(top_radius=6000.0)])
def test_validate_rejects_bad_dims(bad):
    raw = fb.Parameters(**bad).raw()
    assert api._lib().fb_params_validate(ctypes.byref(raw)) == 1
    assert api._lib().fb_last_error()


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fb.FuzzyblueError) as e:
        fb.Builder(0)
    assert e.value.status == 4   # FB_ERR_NO_DEVICE


def test_draw_parameters_are_column_major():
    cols = [[c * 4 + r for r in range(4)] for c in range(4)]
    raw = fb.DrawParameters(cols, [1, 2, 3], [4, 5, 6]).raw()
    flat = np.frombuffer(bytes(raw), dtype=np.float32)
    assert list(flat[:16]) == list(range(16)) and list(flat[16:19]) == [1, 2, 3] and list(flat[20:23]) == [4, 5, 6]
    assert np.array_equal(flat, O.pack_draw(cols, [1, 2, 3], [4, 5, 6]))


def test_header_is_plain_c():
    """include/fuzzyblue.h is the C ABI: it must compile as C99 with no C++ or CUDA headers."""
    import subprocess
    src = '#include "fuzzyblue.h"\nint main(void) { FbParams p; FbDrawParams d; return (int)(sizeof p + sizeof d) == 412 ? 0 : 1; }\n'
    exe = os.path.join(ROOT, "tests", "c_header_check")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe],
                   input=src.encode(), check=True)
    assert subprocess.run([exe]).returncode == 0
    os.remove(exe)
