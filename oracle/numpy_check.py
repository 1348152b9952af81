"""A second, independent restatement of the reference shaders in plain Python / numpy float64 — scalar loops, one
texel at a time, written directly from the GLSL (not from fb_oracle.cpp).

TEST INFRASTRUCTURE ONLY (imported by tests/test_oracle_numpy.py).  Its only job is to catch transcription mistakes in
the C++ oracle: both are run in fp64 on the same unquantised inputs and must agree to rounding.  It is far too slow for
anything but a handful of texels.  Citations: /root/reference/shaders/<file>:<line>.
"""
from __future__ import annotations

import math

import numpy as np

PI = math.pi


class Atm:
    """params.h:26-87 from an oracle.Params (duck-typed)."""

    def __init__(self, p):
        self.p = p
        self.bottom, self.top = float(np.float32(p.bottom_radius)), float(np.float32(p.top_radius))
        f = lambda t: np.array([float(np.float32(v)) for v in t])
        self.solar, self.beta_r, self.beta_m = f(p.solar_irradiance), f(p.rayleigh_scattering), f(p.mie_scattering)
        self.beta_me, self.albedo, self.beta_a = f(p.mie_extinction), f(p.ground_albedo), f(p.absorbtion_extinction)
        self.g, self.mu_s_min = float(np.float32(p.mie_phase_function_g)), float(np.float32(p.mu_s_min))
        self.sun_radius = float(np.float32(p.sun_angular_radius))
        self.H = math.sqrt(self.top ** 2 - self.bottom ** 2)


def clamp(x, lo, hi):
    return min(max(x, lo), hi)


def coord_from_unit(x, n):          # util.h:18-20
    return 0.5 / n + x * (1.0 - 1.0 / n)


def unit_from_coord(u, n):          # util.h:22-24
    return (u - 0.5 / n) / (1.0 - 1.0 / n)


def density(layers, h):             # params.h:89-99
    l = layers[0] if h < float(np.float32(layers[0].width)) else layers[1]
    w = lambda v: float(np.float32(v))
    return clamp(w(l.exp_term) * math.exp(w(l.exp_scale) * h) + w(l.linear_term) * h + w(l.constant_term), 0.0, 1.0)


def dist_top(a, r, mu):             # params.h:105-110
    return max(-r * mu + math.sqrt(max(r * r * (mu * mu - 1.0) + a.top ** 2, 0.0)), 0.0)


def dist_bottom(a, r, mu):          # params.h:112-117
    return max(-r * mu - math.sqrt(max(r * r * (mu * mu - 1.0) + a.bottom ** 2, 0.0)), 0.0)


def hits_ground(a, r, mu):          # params.h:119-124
    return mu < 0.0 and r * r * (mu * mu - 1.0) + a.bottom ** 2 >= 0.0


def rayleigh_phase(nu):             # util.h:26-29
    return 3.0 / (16.0 * PI) * (1.0 + nu * nu)


def mie_phase(g, nu):               # util.h:31-34
    return 3.0 / (8.0 * PI) * (1.0 - g * g) / (2.0 + g * g) * (1.0 + nu * nu) / (1.0 + g * g - 2.0 * g * nu) ** 1.5


def linear_taps(u, n):              # VkSampler LINEAR + CLAMP_TO_EDGE, src/precompute.rs:85-98
    t = u * n - 0.5
    i = math.floor(t)
    return clamp(i, 0, n - 1), clamp(i + 1, 0, n - 1), t - i


def sample2(tab, u, v):             # tab [h][w][4]
    h, w = tab.shape[:2]
    x0, x1, fx = linear_taps(u, w)
    y0, y1, fy = linear_taps(v, h)
    a = tab[y0, x0] * (1 - fx) + tab[y0, x1] * fx
    b = tab[y1, x0] * (1 - fx) + tab[y1, x1] * fx
    return a * (1 - fy) + b * fy


def sample3(tab, u, v, s):          # tab [d][h][w][4]
    d = tab.shape[0]
    z0, z1, fz = linear_taps(s, d)
    return sample2(tab[z0], u, v) * (1 - fz) + sample2(tab[z1], u, v) * fz


# --- transmittance ---------------------------------------------------------------------------------------------
def transmittance_uv(a, r, mu):     # transmittance.h:7-24
    p = a.p
    rho = math.sqrt(max(r * r - a.bottom ** 2, 0.0))
    d, d_min, d_max = dist_top(a, r, mu), a.top - r, rho + a.H
    return coord_from_unit((d - d_min) / (d_max - d_min), p.transmittance_mu_size), coord_from_unit(rho / a.H, p.transmittance_r_size)


def t_to_top(a, T, r, mu):          # transmittance.h:26-33
    return sample2(T, *transmittance_uv(a, r, mu))[:3]


def transmittance(a, T, r, mu, d, hits):   # transmittance.h:35-61
    r_d = clamp(math.sqrt(d * d + 2.0 * r * mu * d + r * r), a.bottom, a.top)
    mu_d = clamp((r * mu + d) / r_d, -1.0, 1.0)
    q = t_to_top(a, T, r_d, -mu_d) / t_to_top(a, T, r, -mu) if hits else t_to_top(a, T, r, mu) / t_to_top(a, T, r_d, mu_d)
    return np.minimum(q, 1.0)


def t_to_sun(a, T, r, mu_s):        # transmittance.h:63-74
    sin_h = a.bottom / r
    cos_h = -math.sqrt(max(1.0 - sin_h * sin_h, 0.0))
    e0, e1 = -sin_h * a.sun_radius, sin_h * a.sun_radius
    t = clamp((mu_s - cos_h - e0) / (e1 - e0), 0.0, 1.0)
    return t_to_top(a, T, r, mu_s) * (t * t * (3.0 - 2.0 * t))


def transmittance_texel(a, x, y):   # transmittance.comp:8-80
    p = a.p
    x_mu, x_r = x / (p.transmittance_mu_size - 1), y / (p.transmittance_r_size - 1)
    rho = a.H * x_r
    r = math.sqrt(rho * rho + a.bottom ** 2)
    d_min, d_max = a.top - r, rho + a.H
    d = d_min + x_mu * (d_max - d_min)
    mu = 1.0 if d == 0.0 else clamp((a.H ** 2 - rho * rho - d * d) / (2.0 * r * d), -1.0, 1.0)

    def optical_length(layers):
        dx = dist_top(a, r, mu) / 500
        s = 0.0
        for i in range(501):
            d_i = i * dx
            r_i = math.sqrt(d_i * d_i + 2.0 * r * mu * d_i + r * r)
            s += density(layers, r_i - a.bottom) * (0.5 if i in (0, 500) else 1.0) * dx
        return s

    tau = a.beta_r * optical_length(p.rayleigh_density) + a.beta_me * optical_length(p.mie_density) + a.beta_a * optical_length(p.absorbtion_density)
    return np.exp(-tau)


# --- scattering texture mappings --------------------------------------------------------------------------------
def scattering_uvwz(a, r, mu, mu_s, nu, hits):   # scattering.h:7-60
    p = a.p
    rho = math.sqrt(max(r * r - a.bottom ** 2, 0.0))
    u_r = coord_from_unit(rho / a.H, p.scattering_r_size)
    r_mu = r * mu
    disc = r_mu * r_mu - r * r + a.bottom ** 2
    half = p.scattering_mu_size // 2
    if hits:
        d, d_min, d_max = -r_mu - math.sqrt(max(disc, 0.0)), r - a.bottom, rho
        u_mu = 0.5 - 0.5 * coord_from_unit(0.0 if d_max == d_min else (d - d_min) / (d_max - d_min), half)
    else:
        d, d_min, d_max = -r_mu + math.sqrt(max(disc + a.H ** 2, 0.0)), a.top - r, rho + a.H
        u_mu = 0.5 + 0.5 * coord_from_unit((d - d_min) / (d_max - d_min), half)
    d = dist_top(a, a.bottom, mu_s)
    d_min, d_max = a.top - a.bottom, a.H
    aa = (d - d_min) / (d_max - d_min)
    A = -2.0 * a.mu_s_min * a.bottom / (d_max - d_min)
    u_mu_s = coord_from_unit(max(1.0 - aa / A, 0.0) / (1.0 + aa), p.scattering_mu_s_size)
    return (nu + 1.0) / 2.0, u_mu_s, u_mu, u_r


def texel_to_geometry(a, x, y, z):  # scattering.h:62-137, util.h:36-45
    p = a.p
    W = p.scattering_nu_size * p.scattering_mu_s_size
    frag = lambda i, n: n * coord_from_unit(i / (n - 1), n)
    fx, fy, fz = frag(x, W), frag(y, p.scattering_mu_size), frag(z, p.scattering_r_size)
    f_nu = math.floor(fx / p.scattering_mu_s_size)
    f_mu_s = fx - p.scattering_mu_s_size * math.floor(fx / p.scattering_mu_s_size)
    u_nu, u_mu_s, u_mu, u_r = f_nu / (p.scattering_nu_size - 1), f_mu_s / p.scattering_mu_s_size, fy / p.scattering_mu_size, fz / p.scattering_r_size
    rho = a.H * unit_from_coord(u_r, p.scattering_r_size)
    r = math.sqrt(rho * rho + a.bottom ** 2)
    half = p.scattering_mu_size // 2
    if u_mu < 0.5:
        d_min, d_max = r - a.bottom, rho
        d = d_min + (d_max - d_min) * unit_from_coord(1.0 - 2.0 * u_mu, half)
        mu = -1.0 if d == 0.0 else clamp(-(rho * rho + d * d) / (2.0 * r * d), -1.0, 1.0)
        hits = True
    else:
        d_min, d_max = a.top - r, rho + a.H
        d = d_min + (d_max - d_min) * unit_from_coord(2.0 * u_mu - 1.0, half)
        mu = 1.0 if d == 0.0 else clamp((a.H ** 2 - rho * rho - d * d) / (2.0 * r * d), -1.0, 1.0)
        hits = False
    x_mu_s = unit_from_coord(u_mu_s, p.scattering_mu_s_size)
    d_min, d_max = a.top - a.bottom, a.H
    A = -2.0 * a.mu_s_min * a.bottom / (d_max - d_min)
    aa = (A - x_mu_s * A) / (1.0 + x_mu_s * A)
    d = d_min + min(aa, A) * (d_max - d_min)
    mu_s = 1.0 if d == 0.0 else clamp((a.H ** 2 - d * d) / (2.0 * a.bottom * d), -1.0, 1.0)
    nu = clamp(u_nu * 2.0 - 1.0, -1.0, 1.0)
    s = math.sqrt((1.0 - mu * mu) * (1.0 - mu_s * mu_s))
    return r, mu, mu_s, clamp(nu, mu * mu_s - s, mu * mu_s + s), hits


def scattering4(a, S, r, mu, mu_s, nu, hits):    # scattering.h:139-155
    n = a.p.scattering_nu_size
    u_nu, u_mu_s, u_mu, u_r = scattering_uvwz(a, r, mu, mu_s, nu, hits)
    tcx = u_nu * (n - 1)
    tx = math.floor(tcx)
    l = tcx - tx
    return sample3(S, (tx + u_mu_s) / n, u_mu, u_r) * (1 - l) + sample3(S, (tx + 1 + u_mu_s) / n, u_mu, u_r) * l


def scattering_order(a, dR, dM, dMS, r, mu, mu_s, nu, hits, order):   # scattering.h:157-179
    if order == 1:
        return scattering4(a, dR, r, mu, mu_s, nu, hits)[:3] * rayleigh_phase(nu) + scattering4(a, dM, r, mu, mu_s, nu, hits)[:3] * mie_phase(a.g, nu)
    return scattering4(a, dMS, r, mu, mu_s, nu, hits)[:3]


# --- the 3-D stages, one texel -----------------------------------------------------------------------------------
def single_scattering_texel(a, T, x, y, z):      # single_scattering.comp:10-65
    p = a.p
    r, mu, mu_s, nu, hits = texel_to_geometry(a, x, y, z)
    dx = (dist_bottom(a, r, mu) if hits else dist_top(a, r, mu)) / 50
    rs, ms = np.zeros(3), np.zeros(3)
    for i in range(51):
        d = i * dx
        r_d = clamp(math.sqrt(d * d + 2.0 * r * mu * d + r * r), a.bottom, a.top)
        mu_s_d = clamp((r * mu_s + d * nu) / r_d, -1.0, 1.0)
        t = transmittance(a, T, r, mu, d, hits) * t_to_sun(a, T, r_d, mu_s_d)
        w = 0.5 if i in (0, 50) else 1.0
        rs += t * density(p.rayleigh_density, r_d - a.bottom) * w
        ms += t * density(p.mie_density, r_d - a.bottom) * w
    return rs * dx * a.solar * a.beta_r, ms * dx * a.solar * a.beta_m


def irradiance_lookup(a, E, r, mu_s):            # irradiance.h:20-38
    p = a.p
    return sample2(E, coord_from_unit(mu_s * 0.5 + 0.5, p.irradiance_mu_s_size),
                   coord_from_unit((r - a.bottom) / (a.top - a.bottom), p.irradiance_r_size))[:3]


def scattering_density_texel(a, T, dR, dM, dMS, dE, x, y, z, order):   # scattering_density.comp:10-107
    p = a.p
    r, mu, mu_s, nu, _ = texel_to_geometry(a, x, y, z)
    omega = np.array([math.sqrt(1.0 - mu * mu), 0.0, mu])
    sx = 0.0 if omega[0] == 0.0 else (nu - mu * mu_s) / omega[0]
    omega_s = np.array([sx, math.sqrt(max(1.0 - sx * sx - mu_s * mu_s, 0.0)), mu_s])
    dphi = dtheta = PI / 16
    acc = np.zeros(3)
    for l in range(16):
        theta = (l + 0.5) * dtheta
        ct, st = math.cos(theta), math.sin(theta)
        hits = hits_ground(a, r, ct)
        dist, t_ground, albedo = 0.0, np.zeros(3), np.zeros(3)
        if hits:
            dist = dist_bottom(a, r, ct)
            t_ground, albedo = transmittance(a, T, r, ct, dist, True), a.albedo
        for m in range(32):
            phi = (m + 0.5) * dphi
            wi = np.array([math.cos(phi) * st, math.sin(phi) * st, ct])
            nu1 = float(omega_s @ wi)
            L = scattering_order(a, dR, dM, dMS, r, wi[2], mu_s, nu1, hits, order - 1)
            gn = np.array([0.0, 0.0, r]) + wi * dist
            gn = gn / math.sqrt(float(gn @ gn))
            L = L + t_ground * albedo * (1.0 / PI) * irradiance_lookup(a, dE, a.bottom, float(gn @ omega_s))
            nu2 = float(omega @ wi)
            h = r - a.bottom
            acc += L * (a.beta_r * density(p.rayleigh_density, h) * rayleigh_phase(nu2)
                        + a.beta_m * density(p.mie_density, h) * mie_phase(a.g, nu2)) * (dtheta * dphi * st)
    return acc


def indirect_irradiance_texel(a, dR, dM, dMS, x, y, order):            # indirect_irradiance.comp:10-74
    p = a.p
    r = a.bottom + y / (p.irradiance_r_size - 1) * (a.top - a.bottom)
    mu_s = clamp(2.0 * x / (p.irradiance_mu_s_size - 1) - 1.0, -1.0, 1.0)
    dphi = dtheta = PI / 32
    omega_s = np.array([math.sqrt(1.0 - mu_s * mu_s), 0.0, mu_s])
    acc = np.zeros(3)
    for j in range(16):
        theta = (j + 0.5) * dtheta
        for i in range(64):
            phi = (i + 0.5) * dphi
            w = np.array([math.cos(phi) * math.sin(theta), math.sin(phi) * math.sin(theta), math.cos(theta)])
            acc += scattering_order(a, dR, dM, dMS, r, w[2], mu_s, float(w @ omega_s), False, order) * w[2] * (dtheta * dphi * math.sin(theta))
    return acc


def multiple_scattering_texel(a, T, dens, x, y, z):                     # multiple_scattering.comp:9-54
    r, mu, mu_s, nu, hits = texel_to_geometry(a, x, y, z)
    dx = (dist_bottom(a, r, mu) if hits else dist_top(a, r, mu)) / 50
    acc = np.zeros(3)
    for i in range(51):
        d = i * dx
        r_i = clamp(math.sqrt(d * d + 2.0 * r * mu * d + r * r), a.bottom, a.top)
        mu_i, mu_s_i = clamp((r * mu + d) / r_i, -1.0, 1.0), clamp((r * mu_s + d * nu) / r_i, -1.0, 1.0)
        acc += scattering4(a, dens, r_i, mu_i, mu_s_i, nu, hits)[:3] * transmittance(a, T, r, mu, d, hits) * dx * (0.5 if i in (0, 50) else 1.0)
    return acc, nu


# --- render_sky.h / render_sky.frag, one pixel ---------------------------------------------------------------------
def extrapolated_single_mie(a, s):               # render_sky.h:9-19
    if s[0] <= 0.0:
        return np.zeros(3)
    return s[:3] * s[3] / s[0] * (a.beta_r[0] / a.beta_m[0]) * (a.beta_m / a.beta_r)


def sky_radiance_to_point(a, T, S, camera, view, point, sun):   # render_sky.h:111-191
    camera, view, point, sun = (np.asarray(v, dtype=np.float64) for v in (camera, view, point, sun))
    r = math.sqrt(float(camera @ camera))
    rmu = float(camera @ view)
    disc = rmu * rmu - r * r + a.top ** 2
    to_top = -rmu - math.sqrt(disc) if disc >= 0 else float("nan")
    if to_top > 0.0:
        camera = camera + view * to_top
        r = a.top
        rmu += to_top
    elif r > a.top:
        return np.zeros(3), np.ones(3)
    mu, mu_s, nu = rmu / r, float(camera @ sun) / r, float(view @ sun)
    d = math.sqrt(float((point - camera) @ (point - camera)))
    hits = hits_ground(a, r, mu)
    tr = transmittance(a, T, r, mu, d, hits)
    sc = scattering4(a, S, r, mu, mu_s, nu, hits)
    mie = extrapolated_single_mie(a, sc)
    sc = sc[:3]
    if not math.isinf(d):
        r_p = clamp(math.sqrt(d * d + 2.0 * r * mu * d + r * r), a.bottom, a.top)
        mu_p, mu_s_p = (r * mu + d) / r_p, (r * mu_s + d * nu) / r_p
        sp = scattering4(a, S, r_p, mu_p, mu_s_p, nu, hits)
        mie_p = extrapolated_single_mie(a, sp)
        sc = sc - tr * sp[:3]
        mie = mie - tr * mie_p
        mie = extrapolated_single_mie(a, np.array([sc[0], sc[1], sc[2], mie[0]]))
        t = clamp(mu_s / 0.01, 0.0, 1.0)
        mie = mie * (t * t * (3.0 - 2.0 * t))
    return sc * rayleigh_phase(nu) + mie * mie_phase(a.g, nu), tr


def render_pixel(a, T, S, draw23, depth, px, py, w, h):          # render_sky.frag:24-35, fullscreen.vert:5-8
    M = np.asarray(draw23[:16], dtype=np.float64).reshape(4, 4).T   # columns are stored contiguously
    cam, sun = np.asarray(draw23[16:19], dtype=np.float64), np.asarray(draw23[20:23], dtype=np.float64)
    ndc = np.array([2.0 * (px + 0.5) / w - 1.0, 2.0 * (py + 0.5) / h - 1.0])
    v = M @ np.array([ndc[0], ndc[1], 0.0, 1.0])
    view = v[:3] / math.sqrt(float(v[:3] @ v[:3]))
    wp = M @ np.array([ndc[0], ndc[1], float(depth), 1.0])
    world = wp[:3] / wp[3] * 1e-3
    return sky_radiance_to_point(a, T, S, cam, view, world, sun)
