"""How many distinct (mu, r) interpolation cells do the 51 trapezoid nodes of one multiple-scattering ray touch?
(multiple_scattering.comp:21-44 through scattering.h:7-41.)  k_multiple_scattering stages one pre-blended table row per
node; nodes of a ray that share a cell could share the four neighbour-row loads.  Pure numpy,
fp64 geometry (cell counts, not bits); prints the distribution over all (r, mu) rows of the default table."""
import numpy as np

bottom, top = 6360.0, 6420.0
R, MU = 32, 128
H = np.sqrt(top * top - bottom * bottom)

def cells_for_row(z, y):
    # texel -> (r, mu), scattering.h:62-118
    rho = H * z / (R - 1)
    r = np.sqrt(rho * rho + bottom * bottom)
    if y < MU // 2:
        d_min, d_max = r - bottom, rho
        x = 1.0 - 2.0 * ((y + 0.5) / MU)
        xm = (x * (MU / 2) - 0.5) / (MU / 2 - 1) if False else (1.0 - 2.0 * (y + 0.5) / MU - 1.0 / MU) / (1.0 - 2.0 / MU)
        d = d_min + (d_max - d_min) * xm
        mu = -1.0 if d == 0 else max(-1.0, min(1.0, -(rho * rho + d * d) / (2 * r * d)))
        hits = True
    else:
        d_min, d_max = top - r, rho + H
        xm = (2.0 * (y + 0.5) / MU - 1.0 - 1.0 / MU) / (1.0 - 2.0 / MU)
        d = d_min + (d_max - d_min) * xm
        mu = 1.0 if d == 0 else max(-1.0, min(1.0, (H * H - rho * rho - d * d) / (2 * r * d)))
        hits = False
    # ray length, params.h:105-133
    if hits:
        L = max(-r * mu - np.sqrt(max(r * r * (mu * mu - 1) + bottom * bottom, 0.0)), 0.0)
    else:
        L = max(-r * mu + np.sqrt(max(r * r * (mu * mu - 1) + top * top, 0.0)), 0.0)
    i = np.arange(51)
    d_i = i * (L / 50.0)
    r_i = np.clip(np.sqrt(d_i * d_i + 2 * r * mu * d_i + r * r), bottom, top)
    mu_i = np.clip((r * mu + d_i) / r_i, -1, 1)
    rho_i = np.sqrt(np.maximum(r_i * r_i - bottom * bottom, 0))
    u_r = 0.5 / R + rho_i / H * (1 - 1.0 / R)
    rmu = r_i * mu_i
    disc = rmu * rmu - r_i * r_i + bottom * bottom
    if hits:
        dd = -rmu - np.sqrt(np.maximum(disc, 0))
        dmin, dmax = r_i - bottom, rho_i
        xm_i = np.where(dmax == dmin, 0.0, (dd - dmin) / np.where(dmax == dmin, 1.0, dmax - dmin))
        u_mu = 0.5 - 0.5 * (0.5 / (MU / 2) + xm_i * (1 - 1.0 / (MU / 2)))
    else:
        dd = -rmu + np.sqrt(np.maximum(disc + H * H, 0))
        dmin, dmax = top - r_i, rho_i + H
        u_mu = 0.5 + 0.5 * (0.5 / (MU / 2) + (dd - dmin) / (dmax - dmin) * (1 - 1.0 / (MU / 2)))
    zc = np.floor(u_r * R - 0.5).astype(int)
    yc = np.floor(u_mu * MU - 0.5).astype(int)
    cells = set(zip(yc.tolist(), zc.tolist()))
    rows = set()
    for (a, b) in cells:
        for da in (0, 1):
            for db in (0, 1):
                rows.add((min(max(a + da, 0), MU - 1), min(max(b + db, 0), R - 1)))
    runs = 1 + int(np.count_nonzero((np.diff(yc) != 0) | (np.diff(zc) != 0)))
    return len(cells), len(rows), runs

tot_cells = tot_rows = tot_runs = 0
hist = np.zeros(52, int)
for z in range(R):
    for y in range(MU):
        c, rws, runs = cells_for_row(z, y)
        tot_cells += c; tot_rows += rws; tot_runs += runs; hist[c] += 1
n = R * MU
print("rows of the table:", n, " nodes per ray: 51 (staged rows now: 4 per node = 204 loads per texel column)")
print("mean distinct (mu, r) cells per ray: %.1f   mean runs of consecutive nodes in one cell: %.1f" % (tot_cells / n, tot_runs / n))
print("mean distinct neighbour rows per ray: %.1f (vs 204 row loads now -> %.2fx fewer loads if each row is loaded once)"
      % (tot_rows / n, 204.0 / (tot_rows / n)))
print("histogram of distinct cells per ray (cells: rays):", {int(k): int(v) for k, v in enumerate(hist) if v})
