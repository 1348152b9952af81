"""Synthetic cameras, depth buffers and randomised atmospheres for tests and bench.py.

The reference defines no render benchmark and no parameter generator; these are the inputs
SURVEY.md §8(d) proposes for BASELINE.json configs 4 and 5.  Host-side numpy only.
"""
from __future__ import annotations

import math
from dataclasses import replace
from typing import List, Tuple

import numpy as np

from .api import DensityProfile, DensityProfileLayer, DrawParameters, Parameters


def _look_at(eye_m: np.ndarray, forward: np.ndarray, up_hint: np.ndarray) -> np.ndarray:
    f = forward / np.linalg.norm(forward)
    s = np.cross(f, up_hint)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = s, u, -f
    view[:3, 3] = -view[:3, :3] @ eye_m
    return view


def _reverse_z_infinite(vfov: float, aspect: float, near: float) -> np.ndarray:
    """Reverse-Z, infinite far plane, Vulkan clip space (y down): depth 1 at `near`, 0 at infinity —
    the convention render_sky.frag:25 assumes (depth 0 is a pure direction)."""
    f = 1.0 / math.tan(vfov / 2)
    p = np.zeros((4, 4))
    p[0, 0] = f / aspect
    p[1, 1] = -f
    p[2, 3] = near
    p[3, 2] = -1.0
    return p


def camera_sweep(n_views: int, width: int, height: int, bottom_radius_km: float = 6360.0, seed: int = 4, altitudes_km=None):
    """`n_views` cameras: altitude log-uniform 1 m .. 2000 km, yaw U[0,2pi), pitch U[-60,60] deg, vfov 60 deg,
    sun zenith angle U[0,110] deg (`altitudes_km`, if given, replaces the random altitudes).
    Returns (list[DrawParameters], list[(M_inv float64, eye_m float64)])."""
    rng = np.random.default_rng(seed)
    draws, extra = [], []
    for k in range(n_views):
        alt_km = math.exp(rng.uniform(math.log(1e-3), math.log(2000.0)))
        if altitudes_km is not None:
            alt_km = float(altitudes_km[k])
        yaw, pitch = rng.uniform(0, 2 * math.pi), math.radians(rng.uniform(-60, 60))
        eye_km = np.array([0.0, 0.0, bottom_radius_km + alt_km])
        fwd = np.array([math.cos(pitch) * math.cos(yaw), math.cos(pitch) * math.sin(yaw), math.sin(pitch)])
        sz, sa = math.radians(rng.uniform(0, 110)), rng.uniform(0, 2 * math.pi)
        sun = np.array([math.sin(sz) * math.cos(sa), math.sin(sz) * math.sin(sa), math.cos(sz)])
        view = _look_at(eye_km * 1e3, fwd, np.array([0.0, 0.0, 1.0]))
        proj = _reverse_z_infinite(math.radians(60), width / height, 0.1)
        inv = np.linalg.inv(proj @ view)
        cols = [[float(np.float32(inv[r, c])) for r in range(4)] for c in range(4)]   # column-major
        draws.append(DrawParameters(cols, [float(v) for v in eye_km], [float(v) for v in sun]))
        extra.append((inv, eye_km * 1e3))
    return draws, extra


def analytic_depth(inv: np.ndarray, eye_m: np.ndarray, width: int, height: int, bottom_radius_km: float = 6360.0,
                   near: float = 0.1) -> np.ndarray:
    """Depth buffer of a bare sphere of radius bottom_radius: reverse-Z depth near/dist_along_forward for pixels
    whose ray hits the ground, 0 (infinitely far: the `isinf(d)` path of render_sky.h:154) for sky pixels."""
    xs = (np.arange(width) + 0.5) / width * 2 - 1
    ys = (np.arange(height) + 0.5) / height * 2 - 1
    nx, ny = np.meshgrid(xs, ys)
    d = inv[:3, 0, None, None] * nx + inv[:3, 1, None, None] * ny + inv[:3, 3, None, None]   # z = 0 row of M^-1
    d = d / np.linalg.norm(d, axis=0)
    R = bottom_radius_km * 1e3
    b = np.einsum("i,ihw->hw", eye_m, d)
    c = eye_m @ eye_m - R * R
    disc = b * b - c
    t = -b - np.sqrt(np.maximum(disc, 0))
    hit = (disc > 0) & (t > 0)
    # view-space z of the hit point: distance along the camera forward axis = t * (d . forward)
    clip_far = inv @ np.array([0.0, 0.0, 0.0, 1.0])        # centre of the screen at depth 0: the forward direction
    fwd = clip_far[:3] / np.linalg.norm(clip_far[:3])
    zview = t * np.einsum("i,ihw->hw", fwd, d)
    depth = np.where(hit, near / np.maximum(zview, 1e-9), 0.0)
    return depth.astype(np.float32)


def random_atmospheres(n: int, base: Parameters = None, seed: int = 20260) -> List[Parameters]:
    """SURVEY.md §8(d) config-4 generator: Earth-to-Mars radii, randomised Rayleigh / Mie / ozone."""
    rng = np.random.default_rng(seed)
    base = base or Parameters()
    out = []
    for _ in range(n):
        bottom = rng.uniform(3389.5, 6360.0)
        top = bottom + rng.uniform(60.0, 120.0)
        ray_scale, ray_h = rng.uniform(0.25, 4.0), rng.uniform(6.0, 12.0)
        mie_s, mie_h, g = rng.uniform(1e-3, 2e-2), rng.uniform(0.8, 2.0), rng.uniform(0.7, 0.9)
        oz = rng.uniform(0.0, 2.0)
        out.append(replace(
            base, bottom_radius=float(bottom), top_radius=float(top),
            rayleigh_scattering=tuple(float(v * ray_scale) for v in (0.005802, 0.013558, 0.033100)),
            rayleigh_density=DensityProfile((DensityProfileLayer(), DensityProfileLayer(0.0, 1.0, float(-1.0 / ray_h), 0.0, 0.0))),
            mie_scattering=(float(mie_s),) * 3, mie_extinction=(float(mie_s / 0.9),) * 3,
            mie_density=DensityProfile((DensityProfileLayer(), DensityProfileLayer(0.0, 1.0, float(-1.0 / mie_h), 0.0, 0.0))),
            mie_phase_function_g=float(g),
            absorbtion_extinction=tuple(float(v * oz) for v in (6.5e-4, 1.881e-3, 8.5e-5)),
            ground_albedo=(float(rng.uniform(0.0, 0.4)),) * 3,
            sun_angular_radius=float(rng.uniform(0.00306, 0.004675))))
    return out
