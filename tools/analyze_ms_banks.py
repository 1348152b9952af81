"""Shared-memory wavefronts of k_multiple_scattering's slab look-ups, simulated on the host (fp64 geometry, default
dims): per (r, mu) row, node and warp, the 16-byte entry index j + kx0 of every lane (scattering.h:43-56 for the
sample's mu_s), and the number of data-pipe wavefronts an LDS.128 needs: 8 lanes per wavefront, distinct addresses in
one 16-byte bank group (entry index mod 8) serialise.  Compares the current layout [node][x] with candidates.
usage: tools/analyze_ms_banks.py [row_stride]   (row_stride: sample every n-th (r, mu) row; default 7)"""
import sys
import numpy as np

bottom, top, mu_s_min = 6360.0, 6420.0, -0.207912
R, MU, MS, NU = 32, 128, 32, 8
W = NU * MS
H = np.sqrt(top * top - bottom * bottom)
stride = int(sys.argv[1]) if len(sys.argv) > 1 and __name__ == "__main__" else 7

def row_geometry(z, y):
    rho = H * z / (R - 1)
    r = np.sqrt(rho * rho + bottom * bottom)
    if y < MU // 2:
        d_min, d_max = r - bottom, rho
        xm = (1.0 - 2.0 * (y + 0.5) / MU - 1.0 / MU) / (1.0 - 2.0 / MU)
        d = d_min + (d_max - d_min) * xm
        mu = -1.0 if d == 0 else max(-1.0, min(1.0, -(rho * rho + d * d) / (2 * r * d)))
        hits = True
        L = max(-r * mu - np.sqrt(max(r * r * (mu * mu - 1) + bottom * bottom, 0.0)), 0.0)
    else:
        d_min, d_max = top - r, rho + H
        xm = (2.0 * (y + 0.5) / MU - 1.0 - 1.0 / MU) / (1.0 - 2.0 / MU)
        d = d_min + (d_max - d_min) * xm
        mu = 1.0 if d == 0 else max(-1.0, min(1.0, (H * H - rho * rho - d * d) / (2 * r * d)))
        hits = False
        L = max(-r * mu + np.sqrt(max(r * r * (mu * mu - 1) + top * top, 0.0)), 0.0)
    return r, mu, L

# texel x -> (mu_s, nu knot), scattering.h:96-118
xs = np.arange(MS) / (MS - 1)
dmin_s, dmax_s = top - bottom, H
A = -2.0 * mu_s_min * bottom / (dmax_s - dmin_s)
a = (A - xs * A) / (1 + xs * A)
d_s = dmin_s + np.minimum(a, A) * (dmax_s - dmin_s)
mu_s_tex = np.where(d_s == 0, 1.0, np.clip((H * H - d_s * d_s) / (2 * bottom * np.where(d_s == 0, 1, d_s)), -1, 1))

def wavefronts(idx):
    """idx: [..., 32] entry indices of one LDS.128 -> wavefronts (4 quarter-warps, conflicts on idx mod 8)."""
    q = idx.reshape(idx.shape[:-1] + (4, 8))
    out = np.zeros(idx.shape[:-1], int)
    for qq in range(4):
        g = q[..., qq, :]
        worst = np.zeros(idx.shape[:-1], int)
        for b in range(8):
            m = (g % 8) == b
            # distinct addresses in this bank group
            vals = np.where(m, g, -1)
            vals = np.sort(vals, axis=-1)
            distinct = (np.diff(vals, axis=-1) != 0).sum(-1) + 1 - (vals[..., 0] == -1)
            worst = np.maximum(worst, distinct)
        out += worst
    return out

def wavefronts_w(idx, lanes, groups):
    """LDS of narrower words: `lanes` lanes per wavefront (16 for LDS.64, 32 for LDS.32), `groups` bank groups."""
    q = idx.reshape(idx.shape[:-1] + (32 // lanes, lanes))
    out = np.zeros(idx.shape[:-1], int)
    for qq in range(32 // lanes):
        g = q[..., qq, :]
        worst = np.zeros(idx.shape[:-1], int)
        for b in range(groups):
            vals = np.sort(np.where((g % groups) == b, g, -1), axis=-1)
            distinct = (np.diff(vals, axis=-1) != 0).sum(-1) + 1 - (vals[..., 0] == -1)
            worst = np.maximum(worst, distinct)
        out += worst
    return out

if __name__ == "__main__":
    tot = {"current": 0, "ideal": 0, "pad1": 0, "pairs32B": 0, "planes_rg_b": 0, "planes_r_g_b": 0}
    n_loads = 0
    i_nodes = np.arange(51)
    for z in range(0, R):
        for y in range(z % stride, MU, stride):
            r, mu, L = row_geometry(z, y)
            d_i = i_nodes * (L / 50.0)
            r_i = np.clip(np.sqrt(d_i * d_i + 2 * r * mu * d_i + r * r), bottom, top)
            for k in range(NU):
                nu_k = 2.0 * k / (NU - 1) - 1.0
                s = np.sqrt(np.maximum((1 - mu * mu) * (1 - mu_s_tex * mu_s_tex), 0))
                nu = np.clip(nu_k, mu * mu_s_tex - s, mu * mu_s_tex + s)                       # [32]
                tcx = (nu + 1) / 2 * (NU - 1)
                tx = np.clip(np.floor(tcx + 1e-6).astype(int), 0, NU - 1)
                mus_i = np.clip((r * mu_s_tex[None, :] + d_i[:, None] * nu[None, :]) / r_i[:, None], -1, 1)   # [51, 32]
                dd = -bottom * mus_i + np.sqrt(bottom * bottom * (mus_i * mus_i - 1) + top * top)
                aa = (dd - dmin_s) / (dmax_s - dmin_s)
                xx = np.maximum(1 - aa / A, 0) / (1 + aa)
                t = np.clip(xx * (MS - 1), 0, MS - 1 - 1e-6)
                j = np.floor(t).astype(int)                                                 # [51, 32]
                e = (i_nodes % 12)[:, None]
                base = e * W + j + (tx * MS)[None, :]
                tot["current"] += wavefronts(base).sum() + wavefronts(base + 1).sum()
                basep = e * (W + 1) + j + (tx * MS)[None, :]
                tot["pad1"] += wavefronts(basep).sum() + wavefronts(basep + 1).sum()
                # (value, next value) pairs in one 32-byte entry: two LDS.128 at 2 * idx and 2 * idx + 1
                tot["pairs32B"] += wavefronts(2 * base).sum() + wavefronts(2 * base + 1).sum()
                # (r, g) float2 plane + b float plane: LDS.64 + LDS.32 per tap, 3 words per lane instead of 4
                tot["planes_rg_b"] += (wavefronts_w(base, 16, 16).sum() + wavefronts_w(base + 1, 16, 16).sum()
                                       + wavefronts_w(base, 32, 32).sum() + wavefronts_w(base + 1, 32, 32).sum())
                tot["planes_r_g_b"] += 3 * (wavefronts_w(base, 32, 32).sum() + wavefronts_w(base + 1, 32, 32).sum())
                tot["ideal"] += 2 * 4 * 51
                n_loads += 2 * 51
    print("LDS.128 per sampled rows:", n_loads, " (two taps per sample, one nu slice)")
    for kname, v in tot.items():
        print("%-11s wavefronts per tap: %.2f" % (kname, v / n_loads))
