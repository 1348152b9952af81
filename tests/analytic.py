"""Closed-form expectations for the 3-D stages of the precompute, in float64 numpy, with NO table look-up and no code
shared with ``oracle/`` or the CUDA library.  Used by tests/test_kat_3d.py to pin the oracle AND the CUDA kernels.

Each expectation feeds a stage a constant (or all-ones) input table, for which the stage's integral collapses to a
formula in the texel's own geometry:

* scattering density, uniform incident radiance c and uniform ground irradiance e, transmittance 1
  (shaders/scattering_density.comp:39-106):
      J(r, mu) = c * sum_lm k(r, mu, l, m) + e * albedo / pi * sum_{l: theta_l reaches the ground} sum_m k(r, mu, l, m)
      k = (beta_R rho_R(r) P_R(nu2) + beta_M rho_M(r) P_M(nu2)) sin(theta_l) dtheta dphi,  nu2 = omega(mu) . omega_i(l, m)
  (phase functions integrate to 1: the first sum is within 1e-3 of c (beta_R rho_R + beta_M rho_M)).
* indirect irradiance, uniform radiance c (shaders/indirect_irradiance.comp:27-44):
      E = c * sum_j sum_i cos(theta_j) sin(theta_j) dtheta dphi  ->  pi c  (16 x 64 midpoint rule: pi c (1 + 1.6e-3))
* multiple scattering, uniform density J, transmittance 1 (shaders/multiple_scattering.comp:22-52):
      delta = J * (0.5 + 49 + 0.5) dx = J * d(texel), d = distance to the nearest boundary, which the texel mapping of
      shaders/scattering.h:84-107 gives as a LINEAR function of the texel's mu coordinate; scattering += delta / P_R(nu).
* single scattering, transmittance 1 (shaders/single_scattering.comp:10-70, transmittance.h:63-74):
      rayleigh = E_sun beta_R dx sum_i w_i rho_R(r_i) vis(r_i, mu_s_i), vis = the smoothstep of the sun disc at the
      local horizon; where the sun stays above the horizon all along the ray this is the trapezoid rule of
      int rho_R ds, checked against scipy.integrate.quad.

The texel -> (r, mu, mu_s, nu) mapping below restates shaders/scattering.h:62-137 (Bruneton's mapping with the older `A`
formula, integer mu_size / 2) in float64.
"""
from __future__ import annotations

import math

import numpy as np


def _unit_from_coord(u, n):          # util.h:22-24
    return (u - 0.5 / n) / (1.0 - 1.0 / n)


def profile_density(layers, h):      # params.h:89-99
    h = np.asarray(h, dtype=np.float64)
    def lay(l):
        return np.clip(l.exp_term * np.exp(l.exp_scale * h) + l.linear_term * h + l.constant_term, 0.0, 1.0)
    return np.where(h < layers[0].width, lay(layers[0]), lay(layers[1]))


def rayleigh_phase(nu):              # util.h:26-29
    return 3.0 / (16.0 * math.pi) * (1.0 + nu * nu)


def mie_phase(g, nu):                # util.h:31-34
    return 3.0 / (8.0 * math.pi) * (1.0 - g * g) / (2.0 + g * g) * (1.0 + nu * nu) / (1.0 + g * g - 2.0 * g * nu) ** 1.5


class Geometry:
    """Per-texel geometry of the scattering table, arrays broadcastable to [R][MU][NU][MS]."""

    def __init__(self, p):
        NR, NMU, NMS, NNU = p.scattering_r_size, p.scattering_mu_size, p.scattering_mu_s_size, p.scattering_nu_size
        bot, top = float(np.float32(p.bottom_radius)), float(np.float32(p.top_radius))
        self.shape = (NR, NMU, NNU, NMS)
        self.bot, self.top = bot, top
        H = math.sqrt(top * top - bot * bot)
        z = np.arange(NR, dtype=np.float64)[:, None, None, None]
        y = np.arange(NMU, dtype=np.float64)[None, :, None, None]
        k = np.arange(NNU, dtype=np.float64)[None, None, :, None]
        ms = np.arange(NMS, dtype=np.float64)[None, None, None, :]
        rho = H * _unit_from_coord((z + 0.5) / NR, NR)
        self.rho = rho
        self.r = np.sqrt(rho * rho + bot * bot)
        r = self.r
        u_mu = (y + 0.5) / NMU
        self.hits = np.broadcast_to(u_mu < 0.5, (NR, NMU, 1, 1))
        half = NMU // 2
        with np.errstate(all="ignore"):
            d_g = (r - bot) + (rho - (r - bot)) * _unit_from_coord(1.0 - 2.0 * u_mu, half)
            mu_g = np.where(d_g == 0, -1.0, np.clip(-(rho * rho + d_g * d_g) / (2 * r * d_g), -1, 1))
            d_t = (top - r) + (rho + H - (top - r)) * _unit_from_coord(2.0 * u_mu - 1.0, half)
            mu_t = np.where(d_t == 0, 1.0, np.clip((H * H - rho * rho - d_t * d_t) / (2 * r * d_t), -1, 1))
        self.mu = np.where(u_mu < 0.5, mu_g, mu_t)
        # distance to the nearest boundary along the texel's ray: the mapping variable itself (scattering.h:87-90, :98-101)
        self.d = np.where(u_mu < 0.5, d_g, d_t)
        mu_s_min = float(np.float32(p.mu_s_min))
        d_min, d_max = top - bot, H
        A = -2.0 * mu_s_min * bot / (d_max - d_min)
        x = _unit_from_coord((ms + 0.5) / NMS, NMS)
        a = (A - x * A) / (1.0 + x * A)
        dd = d_min + np.minimum(a, A) * (d_max - d_min)
        self.mu_s = np.where(dd == 0, 1.0, np.clip((H * H - dd * dd) / (2 * bot * dd), -1, 1))
        nu = np.clip(k / (NNU - 1) * 2.0 - 1.0, -1, 1)
        s = np.sqrt((1 - self.mu ** 2) * (1 - self.mu_s ** 2))
        self.nu = np.clip(nu, self.mu * self.mu_s - s, self.mu * self.mu_s + s)       # scattering.h:135-136

    def table(self, a):
        """[R][MU][NU][MS] (+ trailing channel axis) -> the linear [R][MU][NU*MS] layout of the images."""
        a = np.broadcast_to(a, self.shape + a.shape[4:]) if a.ndim > 4 else np.broadcast_to(a, self.shape)
        return a.reshape(self.shape[0], self.shape[1], self.shape[2] * self.shape[3], *a.shape[4:])


def _f32(v):
    return np.asarray(np.asarray(v, dtype=np.float32), dtype=np.float64)


def density_uniform(p, c_rgb, e_rgb):
    """scattering_density for incident radiance == c_rgb everywhere, ground irradiance == e_rgb, transmittance == 1.
    Returns [R][MU][NU*MS][3]."""
    g = Geometry(p)
    N = 16
    dth = dph = math.pi / N
    th = (np.arange(N) + 0.5) * dth
    ph = (np.arange(2 * N) + 0.5) * dph
    ct, st = np.cos(th)[:, None], np.sin(th)[:, None]
    cp = np.cos(ph)[None, :]
    r, mu = g.r[:, :, 0, 0], g.mu[:, :, 0, 0]                      # [R][1] and [R][MU]
    ox = np.sqrt(1.0 - mu * mu)
    nu2 = ox[..., None, None] * (cp * st) + mu[..., None, None] * ct        # [R][MU][16][32]
    dw = dth * dph * st
    gM = float(np.float32(p.mie_phase_function_g))
    PR = (rayleigh_phase(nu2) * dw)                                 # [R][MU][16][32]
    PM = (mie_phase(gM, nu2) * dw)
    # rows whose ray reaches the ground (params.h:119-124), per (r, l)
    rr = g.r[:, 0, 0, 0][:, None]
    hits = (np.cos(th)[None, :] < 0) & (rr * rr * (np.cos(th)[None, :] ** 2 - 1.0) + g.bot * g.bot >= 0)     # [R][16]
    hm = hits[:, None, :, None]
    sums = {"R": PR.sum((-1, -2)), "M": PM.sum((-1, -2)), "Rg": (PR * hm).sum((-1, -2)), "Mg": (PM * hm).sum((-1, -2))}
    h = (r - g.bot)
    rhoR = profile_density(p.rayleigh_density, h)                   # [R][1]
    rhoM = profile_density(p.mie_density, h)
    bR, bM = _f32(p.rayleigh_scattering), _f32(p.mie_scattering)
    alb = _f32(p.ground_albedo)
    c, e = np.asarray(c_rgb, dtype=np.float64), np.asarray(e_rgb, dtype=np.float64)
    kall = bR * (rhoR * sums["R"])[..., None] + bM * (rhoM * sums["M"])[..., None]          # [R][MU][3]
    kgnd = bR * (rhoR * sums["Rg"])[..., None] + bM * (rhoM * sums["Mg"])[..., None]
    out = c * kall + e * alb / math.pi * kgnd
    W = p.scattering_nu_size * p.scattering_mu_s_size
    return np.broadcast_to(out[:, :, None, :], (out.shape[0], out.shape[1], W, 3)).copy(), sums


def indirect_irradiance_uniform(c_rgb):
    """delta_irradiance for radiance == c_rgb everywhere: the 16 x 64 midpoint sum of cos(theta) over the hemisphere."""
    N = 32
    dth = dph = math.pi / N
    th = (np.arange(N // 2) + 0.5) * dth
    s = float((np.cos(th) * np.sin(th)).sum() * dth * (2 * N) * dph)
    return np.asarray(c_rgb, dtype=np.float64) * s, s


def multiple_scattering_uniform(p, J_rgb):
    """(delta_multiple_scattering, scattering increment) for density == J_rgb everywhere and transmittance == 1."""
    g = Geometry(p)
    J = np.asarray(J_rgb, dtype=np.float64)
    delta = g.d[..., None] * J                                      # [R][MU][1][1][3]
    inc = delta / rayleigh_phase(g.nu)[..., None]                   # [R][MU][NU][MS][3]
    return g.table(np.broadcast_to(delta, g.shape + (3,))), g.table(inc), g


def single_scattering_unit_transmittance(p):
    """(delta_rayleigh, delta_mie) [R][MU][NU*MS][3] for a transmittance table == 1: the 51-point trapezoid of
    rho(r_d) * sun-disc visibility along each texel's ray."""
    g = Geometry(p)
    n = 50
    dx = g.d / n                                                    # [R][MU][1][1]
    w = np.ones(n + 1)
    w[0] = w[-1] = 0.5
    alpha = float(np.float32(p.sun_angular_radius))
    sumR = np.zeros(g.shape)
    sumM = np.zeros(g.shape)
    for i in range(n + 1):
        d = i * dx
        r_d = np.clip(np.sqrt(d * d + 2.0 * g.r * g.mu * d + g.r * g.r), g.bot, g.top)
        mu_s_d = np.clip((g.r * g.mu_s + d * g.nu) / r_d, -1, 1)
        sh = g.bot / r_d
        ch = -np.sqrt(np.maximum(1.0 - sh * sh, 0.0))
        t = np.clip((mu_s_d - ch + sh * alpha) / (2 * sh * alpha), 0.0, 1.0)
        vis = t * t * (3.0 - 2.0 * t)                               # smoothstep
        h = r_d - g.bot
        sumR += w[i] * profile_density(p.rayleigh_density, h) * vis
        sumM += w[i] * profile_density(p.mie_density, h) * vis
    sol = _f32(p.solar_irradiance)
    ray = (sumR * dx)[..., None] * (sol * _f32(p.rayleigh_scattering))
    mie = (sumM * dx)[..., None] * (sol * _f32(p.mie_scattering))
    return g.table(ray), g.table(mie), g
