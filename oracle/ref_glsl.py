"""ctypes front-end of oracle/_ref/libfb_glsl_ref.so: the REFERENCE'S OWN SHADERS (/root/reference/shaders/*.h, *.comp,
render_sky.frag) compiled as C++ by oracle/glsl_ref/ and run on the CPU.

TEST INFRASTRUCTURE ONLY.  It exists to pin the oracle (oracle/fb_oracle.cpp, a restatement) against the reference
itself, and to generate the golden vectors of tests/golden/ref_glsl_*.npz.  The library is built only where
/root/reference is present (this container); the built .so travels to the GPU box, the sources never enter the
repository.  Same call signatures as oracle/oracle.py (arithmetic mode fixed: fp32 as written).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfb_glsl_ref.so")
REFERENCE_SHADERS = "/root/reference/shaders"
_lib = None


def build() -> bool:
    """make -C oracle/glsl_ref (needs the reference checkout); False if the reference is not present here."""
    if not os.path.isdir(REFERENCE_SHADERS):
        return os.path.exists(LIB_PATH)
    subprocess.run(["make", "-C", os.path.join(_HERE, "glsl_ref"), "-s", "-j4"], check=True)
    return True


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(LIB_PATH)
        d, f, i64 = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        vp, ci, c64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        L.fbr_transmittance.argtypes = [vp, d]
        L.fbr_direct_irradiance.argtypes = [vp, d, d]
        L.fbr_single_scattering.argtypes = [vp, d, i64, c64, d, d, d]
        L.fbr_scattering_density.argtypes = [vp, ci, d, d, d, d, d, i64, c64, d]
        L.fbr_indirect_irradiance.argtypes = [vp, ci, d, d, d, d, d]
        L.fbr_multiple_scattering.argtypes = [vp, d, d, i64, c64, d, d]
        L.fbr_render.argtypes = [vp, d, d, f, f, ci, ci, i64, c64, d, d]
        L.fbr_set_threads.argtypes = [ci]
        L.fbr_max_threads.restype = ci
        _lib = L
    return _lib


_p, _c, _idx = O._p, O._c, O._idx


def set_threads(n: int) -> int:
    lib().fbr_set_threads(int(n))
    return int(lib().fbr_max_threads())


def transmittance(p: O.Params) -> np.ndarray:
    T = np.zeros(p.t_shape)
    assert lib().fbr_transmittance(p.pack(), _p(T)) == 0
    return T


def direct_irradiance(p: O.Params, T) -> np.ndarray:
    dE, T = np.zeros(p.e_shape), _c(T)
    assert lib().fbr_direct_irradiance(p.pack(), _p(T), _p(dE)) == 0
    return dE


def single_scattering(p: O.Params, T, idx=None):
    T = _c(T)
    ip, n, keep = _idx(idx)
    shape = p.s_shape if idx is None else (n, 4)
    dR, dM, S = np.zeros(shape), np.zeros(shape), np.zeros(shape)
    assert lib().fbr_single_scattering(p.pack(), _p(T), ip, n, _p(dR), _p(dM), _p(S)) == 0
    return dR, dM, S


def scattering_density(p: O.Params, order: int, T, dR, dM, dMS, dE, idx=None) -> np.ndarray:
    T, dR, dM, dMS, dE = map(_c, (T, dR, dM, dMS, dE))
    ip, n, keep = _idx(idx)
    out = np.zeros(p.s_shape if idx is None else (n, 4))
    assert lib().fbr_scattering_density(p.pack(), order, _p(T), _p(dR), _p(dM), _p(dMS), _p(dE), ip, n, _p(out)) == 0
    return out


def indirect_irradiance(p: O.Params, order: int, dR, dM, dMS, E):
    dR, dM, dMS = map(_c, (dR, dM, dMS))
    E = _c(E).copy()
    dE = np.zeros(p.e_shape)
    assert lib().fbr_indirect_irradiance(p.pack(), order, _p(dR), _p(dM), _p(dMS), _p(dE), _p(E)) == 0
    return dE, E


def multiple_scattering(p: O.Params, T, dens, S, idx=None):
    T, dens = _c(T), _c(dens)
    ip, n, keep = _idx(idx)
    if idx is None:
        S = _c(S).copy()
        dMS = np.zeros(p.s_shape)
    else:
        S = _c(_c(S).reshape(-1, 4)[keep]).copy()
        dMS = np.zeros((n, 4))
    assert lib().fbr_multiple_scattering(p.pack(), _p(T), _p(dens), ip, n, _p(dMS), _p(S)) == 0
    return dMS, S


def render(p: O.Params, T, S, draw: np.ndarray, depth: np.ndarray, idx=None):
    T, S = _c(T), _c(S)
    h, w = depth.shape
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    draw = np.ascontiguousarray(draw, dtype=np.float32)
    ip, n, keep = _idx(idx)
    shape = (h, w, 4) if idx is None else (n, 4)
    color, transm = np.zeros(shape), np.zeros(shape)
    fp = ctypes.POINTER(ctypes.c_float)
    assert lib().fbr_render(p.pack(), _p(T), _p(S), draw.ctypes.data_as(fp), depth.ctypes.data_as(fp), w, h, ip, n, _p(color), _p(transm)) == 0
    return color, transm


def precompute(p: O.Params, keep_history: bool = False) -> O.Tables:
    """The recorded command stream of Atmosphere::build (src/precompute.rs:1671-2048) over the reference's shaders."""
    T = transmittance(p)
    dE = direct_irradiance(p, T)
    dR, dM, S = single_scattering(p, T)
    E = np.zeros(p.e_shape)
    dMS = np.zeros(p.s_shape)
    dens, hist = None, {}
    if keep_history:
        hist["single"] = dict(delta_irradiance=dE.copy(), scattering=S.copy())
    for order in range(2, p.order + 1):
        dens = scattering_density(p, order, T, dR, dM, dMS, dE)
        dE, E = indirect_irradiance(p, order - 1, dR, dM, dMS, E)
        dMS, S = multiple_scattering(p, T, dens, S)
        if keep_history:
            hist[order] = dict(scattering_density=dens.copy(), delta_irradiance=dE.copy(), irradiance=E.copy(),
                               delta_multiple_scattering=dMS.copy(), scattering=S.copy())
    return O.Tables(T, E, S, dE, dR, dM, dMS if p.order >= 2 else None, dens, hist)
