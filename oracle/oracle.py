"""ctypes front-end of the CPU oracle (``fb_oracle.cpp``) + the reference's precompute schedule.

TEST INFRASTRUCTURE ONLY — imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never by ``fuzzyblue_b200``.
PARITY PIN (see the header of ``fb_oracle.cpp``): mode 0 reproduces, bit for bit, the reference's own shaders compiled
as C++ and run on the CPU (``oracle/ref_glsl.py``) and the golden vectors generated from them (``tests/golden/``).

Modes: 0 = fp32 arithmetic, reference storage formats ("the shaders as written");
       1 = fp64 arithmetic, reference storage formats; 2 = fp64, no quantisation.
All tables are float64 numpy arrays shaped [r][mu][nu*mu_s][4] / [h][w][4], i.e. the linear
read-back layout of /root/reference/examples/dump.rs:175-193.
"""
from __future__ import annotations

import ctypes
import os
import struct
import subprocess
from dataclasses import dataclass, field, replace
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfb_oracle.so")
_lib = None

F32, F64Q, F64 = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile libfb_oracle.so with oracle/Makefile (g++, OpenMP, -ffp-contract=off)."""
    src = os.path.join(_HERE, "fb_oracle.cpp")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def set_threads(n: int) -> int:
    """Number of OpenMP worker threads of the texel loops (returns the count in effect)."""
    lib().fbo_set_threads(int(n))
    return int(lib().fbo_max_threads())


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        d = ctypes.POINTER(ctypes.c_double)
        f = ctypes.POINTER(ctypes.c_float)
        i64 = ctypes.POINTER(ctypes.c_int64)
        vp, ci, c64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        L.fbo_params_size.restype = ci
        L.fbo_set_threads.argtypes = [ci]
        L.fbo_max_threads.restype = ci
        L.fbo_round_to_half.restype = ctypes.c_double
        L.fbo_round_to_half.argtypes = [ctypes.c_double]
        L.fbo_transmittance.argtypes = [vp, ci, d]
        L.fbo_direct_irradiance.argtypes = [vp, ci, d, d]
        L.fbo_single_scattering.argtypes = [vp, ci, d, i64, c64, d, d, d]
        L.fbo_scattering_density.argtypes = [vp, ci, ci, d, d, d, d, d, i64, c64, d]
        L.fbo_indirect_irradiance.argtypes = [vp, ci, ci, d, d, d, d, d]
        L.fbo_multiple_scattering.argtypes = [vp, ci, d, d, i64, c64, d, d]
        L.fbo_render.argtypes = [vp, ci, d, d, f, f, ci, ci, i64, c64, d, d]
        L.fbo_sky_radiance.argtypes = [vp, ci, d, d, d, d, d, c64, d, d]
        L.fbo_sun_sky_irradiance.argtypes = [vp, ci, d, d, d, d, d, c64, d, d]
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------
# Parameters (src/precompute.rs:690-769, defaults :849-935) and their 320-byte block (:937-1033)
# ---------------------------------------------------------------------------------------------
@dataclass
class Layer:
    width: float = 0.0
    exp_term: float = 0.0
    exp_scale: float = 0.0
    linear_term: float = 0.0
    constant_term: float = 0.0


@dataclass
class Params:
    order: int = 4
    transmittance_mu_size: int = 256
    transmittance_r_size: int = 64
    scattering_r_size: int = 32
    scattering_mu_size: int = 128
    scattering_mu_s_size: int = 32
    scattering_nu_size: int = 8
    irradiance_mu_s_size: int = 64
    irradiance_r_size: int = 16
    solar_irradiance: tuple = (1.474, 1.850, 1.91198)
    sun_angular_radius: float = 0.004675
    bottom_radius: float = 6360.0
    top_radius: float = 6420.0
    rayleigh_density: tuple = (Layer(), Layer(0.0, 1.0, -0.125, 0.0, 0.0))
    rayleigh_scattering: tuple = (0.005802, 0.013558, 0.033100)
    mie_density: tuple = (Layer(), Layer(0.0, 1.0, -0.833333, 0.0, 0.0))
    mie_scattering: tuple = (0.003996, 0.003996, 0.003996)
    mie_extinction: tuple = (0.004440, 0.004440, 0.004440)
    mie_phase_function_g: float = 0.8
    absorbtion_density: tuple = (Layer(25.0, 0.0, 0.0, 0.066667, -0.666667),
                                 Layer(0.0, 0.0, 0.0, -0.066667, 2.666667))
    absorbtion_extinction: tuple = (6.5e-4, 1.881e-3, 8.5e-5)
    ground_albedo: tuple = (0.1, 0.1, 0.1)
    mu_s_min: float = -0.207912

    def pack(self) -> bytes:
        b = struct.pack("<3ff3ff3ff3ff3ff3f", *self.solar_irradiance, self.sun_angular_radius,
                        *self.rayleigh_scattering, self.bottom_radius,
                        *self.mie_scattering, self.top_radius,
                        *self.mie_extinction, self.mie_phase_function_g,
                        *self.ground_albedo, self.mu_s_min,
                        *self.absorbtion_extinction)
        b += struct.pack("<8i", self.transmittance_mu_size, self.transmittance_r_size,
                         self.scattering_r_size, self.scattering_mu_size, self.scattering_mu_s_size,
                         self.scattering_nu_size, self.irradiance_mu_s_size, self.irradiance_r_size)
        b += b"\0" * 4
        for prof in (self.rayleigh_density, self.mie_density, self.absorbtion_density):
            for l in prof:
                b += struct.pack("<5f12x", l.width, l.exp_term, l.exp_scale, l.linear_term, l.constant_term)
        assert len(b) == 320
        return b

    # extents, src/precompute.rs:771-793
    @property
    def t_shape(self):
        return (self.transmittance_r_size, self.transmittance_mu_size, 4)

    @property
    def e_shape(self):
        return (self.irradiance_r_size, self.irradiance_mu_s_size, 4)

    @property
    def s_shape(self):
        return (self.scattering_r_size, self.scattering_mu_size,
                self.scattering_nu_size * self.scattering_mu_s_size, 4)

    def small(self, **kw) -> "Params":
        return replace(self, **kw)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _idx(idx):
    if idx is None:
        return None, 0, None
    a = np.ascontiguousarray(idx, dtype=np.int64)
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), a.size, a


def round_to_half(x):
    L = lib()
    return np.vectorize(lambda v: L.fbo_round_to_half(float(v)))(np.asarray(x, dtype=np.float64))


def transmittance(p: Params, mode: int) -> np.ndarray:
    T = np.zeros(p.t_shape)
    assert lib().fbo_transmittance(p.pack(), mode, _p(T)) == 0
    return T


def direct_irradiance(p: Params, mode: int, T) -> np.ndarray:
    dE = np.zeros(p.e_shape)
    T = _c(T)
    assert lib().fbo_direct_irradiance(p.pack(), mode, _p(T), _p(dE)) == 0
    return dE


def single_scattering(p: Params, mode: int, T, idx=None):
    T = _c(T)
    ip, n, keep = _idx(idx)
    shape = p.s_shape if idx is None else (n, 4)
    dR, dM, S = np.zeros(shape), np.zeros(shape), np.zeros(shape)
    assert lib().fbo_single_scattering(p.pack(), mode, _p(T), ip, n, _p(dR), _p(dM), _p(S)) == 0
    return dR, dM, S


def scattering_density(p: Params, mode: int, order: int, T, dR, dM, dMS, dE, idx=None) -> np.ndarray:
    """``order`` is the push constant of src/precompute.rs:1897 (2 on the first loop pass)."""
    T, dR, dM, dMS, dE = map(_c, (T, dR, dM, dMS, dE))
    ip, n, keep = _idx(idx)
    out = np.zeros(p.s_shape if idx is None else (n, 4))
    assert lib().fbo_scattering_density(p.pack(), mode, order, _p(T), _p(dR), _p(dM), _p(dMS), _p(dE),
                                        ip, n, _p(out)) == 0
    return out


def indirect_irradiance(p: Params, mode: int, order: int, dR, dM, dMS, E):
    """``order`` is the push constant of src/precompute.rs:1946 (loop order - 1).
    Returns (delta_irradiance, irradiance + delta)."""
    dR, dM, dMS = map(_c, (dR, dM, dMS))
    E = _c(E).copy()
    dE = np.zeros(p.e_shape)
    assert lib().fbo_indirect_irradiance(p.pack(), mode, order, _p(dR), _p(dM), _p(dMS), _p(dE), _p(E)) == 0
    return dE, E


def multiple_scattering(p: Params, mode: int, T, dens, S, idx=None):
    """Returns (delta_multiple_scattering, scattering accumulated) — whole tables, or [n][4] rows
    for the linear texel indices ``idx`` (S rows are gathered from the S table given)."""
    T, dens = _c(T), _c(dens)
    ip, n, keep = _idx(idx)
    if idx is None:
        S = _c(S).copy()
        dMS = np.zeros(p.s_shape)
    else:
        S = _c(_c(S).reshape(-1, 4)[keep]).copy()
        dMS = np.zeros((n, 4))
    assert lib().fbo_multiple_scattering(p.pack(), mode, _p(T), _p(dens), ip, n, _p(dMS), _p(S)) == 0
    return dMS, S


def pack_draw(inverse_viewproj, camera_position, sun_direction) -> np.ndarray:
    """DrawParamsRaw, src/render.rs:254-271: 92 bytes = 23 floats; ``inverse_viewproj`` is
    [[f32;4];4] with each inner array one *column* (GLSL mat4 is column-major)."""
    a = np.zeros(23, dtype=np.float32)
    a[:16] = np.asarray(inverse_viewproj, dtype=np.float32).reshape(16)
    a[16:19] = camera_position
    a[20:23] = sun_direction
    return a


def render(p: Params, mode: int, T, S, draw: np.ndarray, depth: np.ndarray, idx=None):
    T, S = _c(T), _c(S)
    h, w = depth.shape
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    draw = np.ascontiguousarray(draw, dtype=np.float32)
    ip, n, keep = _idx(idx)
    shape = (h, w, 4) if idx is None else (n, 4)
    color, transm = np.zeros(shape), np.zeros(shape)
    fp = ctypes.POINTER(ctypes.c_float)
    assert lib().fbo_render(p.pack(), mode, _p(T), _p(S), draw.ctypes.data_as(fp), depth.ctypes.data_as(fp),
                            w, h, ip, n, _p(color), _p(transm)) == 0
    return color, transm


def sky_radiance(p: Params, mode: int, T, S, cam, view, sun):
    T, S, cam, view, sun = map(_c, (T, S, cam, view, sun))
    n = cam.shape[0]
    rad, tr = np.zeros((n, 3)), np.zeros((n, 3))
    assert lib().fbo_sky_radiance(p.pack(), mode, _p(T), _p(S), _p(cam), _p(view), _p(sun), n, _p(rad), _p(tr)) == 0
    return rad, tr


def sun_sky_irradiance(p: Params, mode: int, T, E, point, normal, sun):
    T, E, point, normal, sun = map(_c, (T, E, point, normal, sun))
    n = point.shape[0]
    direct, sky = np.zeros((n, 3)), np.zeros((n, 3))
    assert lib().fbo_sun_sky_irradiance(p.pack(), mode, _p(T), _p(E), _p(point), _p(normal), _p(sun), n,
                                        _p(direct), _p(sky)) == 0
    return direct, sky


@dataclass
class Tables:
    transmittance: np.ndarray
    irradiance: np.ndarray
    scattering: np.ndarray
    delta_irradiance: np.ndarray
    delta_rayleigh: np.ndarray
    delta_mie: np.ndarray
    delta_multiple_scattering: Optional[np.ndarray] = None
    scattering_density: Optional[np.ndarray] = None
    history: dict = field(default_factory=dict)


def precompute(p: Params, mode: int, keep_history: bool = False) -> Tables:
    """The recorded command stream of Atmosphere::build, src/precompute.rs:1671-2048:
    K1 transmittance, K2 direct irradiance -> delta_irradiance, K3 single scattering,
    irradiance cleared to 0 (direct irradiance is *not* accumulated, :1802-1831), then for
    order in 2..=p.order: K4 density(order) reading the previous pass's delta_irradiance,
    K5 indirect irradiance(order-1), K6 multiple scattering."""
    T = transmittance(p, mode)
    dE = direct_irradiance(p, mode, T)
    dR, dM, S = single_scattering(p, mode, T)
    E = np.zeros(p.e_shape)
    dMS = np.zeros(p.s_shape)   # bound but never read at order 2 (scattering.h:165-178)
    dens = None
    hist = {}
    if keep_history:
        hist["single"] = dict(delta_irradiance=dE.copy(), scattering=S.copy())
    for order in range(2, p.order + 1):
        dens = scattering_density(p, mode, order, T, dR, dM, dMS, dE)
        dE, E = indirect_irradiance(p, mode, order - 1, dR, dM, dMS, E)
        dMS, S = multiple_scattering(p, mode, T, dens, S)
        if keep_history:
            hist[order] = dict(scattering_density=dens.copy(), delta_irradiance=dE.copy(),
                               irradiance=E.copy(), delta_multiple_scattering=dMS.copy(),
                               scattering=S.copy())
    return Tables(T, E, S, dE, dR, dM, dMS if p.order >= 2 else None, dens, hist)
