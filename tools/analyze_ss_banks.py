"""Shared-memory wavefronts of k_single_scattering's two slab taps (the pre-blended transmittance row of a node at the
column of GetTransmittanceToSun(r_d, mu_s_d), transmittance.h:7-24,63-74), replayed on the host in fp64 for the default
dims.  Compares the product thread -> texel mapping (a warp = the 32 mu_s values of one nu slice) with a transposed one
(a quarter-warp = the 8 nu slices of one mu_s, whose columns nearly coincide).
usage: tools/analyze_ss_banks.py [row_stride]"""
import sys
import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from analyze_ms_banks import row_geometry, wavefronts, mu_s_tex, bottom, top, H, R, MU, MS, NU, W  # noqa: E402

TW = 256
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 13
tot = {"product": 0, "transposed": 0, "ideal": 0}
n = 0
i_nodes = np.arange(51)
for z in range(R):
    for y in range(z % stride, MU, stride):
        r, mu, L = row_geometry(z, y)
        d_i = i_nodes * (L / 50.0)
        r_d = np.clip(np.sqrt(d_i * d_i + 2 * r * mu * d_i + r * r), bottom, top)
        rho = np.sqrt(np.maximum(r_d * r_d - bottom * bottom, 0))
        d_min, d_max = top - r_d, rho + H
        j_all = np.zeros((51, NU, MS), int)
        for k in range(NU):
            nu_k = 2.0 * k / (NU - 1) - 1.0
            s = np.sqrt(np.maximum((1 - mu * mu) * (1 - mu_s_tex * mu_s_tex), 0))
            nu = np.clip(nu_k, mu * mu_s_tex - s, mu * mu_s_tex + s)
            mu_s_d = np.clip((r * mu_s_tex[None, :] + d_i[:, None] * nu[None, :]) / r_d[:, None], -1, 1)
            disc = r_d[:, None] ** 2 * (mu_s_d ** 2 - 1) + top * top
            dtop = np.maximum(-r_d[:, None] * mu_s_d + np.sqrt(np.maximum(disc, 0)), 0)
            tu = np.clip((dtop - d_min[:, None]) / (d_max - d_min)[:, None] * (TW - 1), 0, TW - 1 - 1e-4)
            j_all[:, k, :] = np.floor(tu).astype(int)
        e = (i_nodes % 12)[:, None]
        # product mapping: warp k holds the 32 mu_s of nu slice k
        for k in range(NU):
            idx = e * TW + j_all[:, k, :]
            tot["product"] += wavefronts(idx).sum() + wavefronts(idx + 1).sum()
        # transposed: thread t -> (nu = t % NU, mu_s = t // NU); warp w holds mu_s 4w .. 4w+3, each with its 8 nu slices
        for w in range(MS * NU // 32):
            cols = j_all[:, :, 4 * w:4 * w + 4]                       # [51, nu 8, mu_s 4]
            idx = e * TW + np.transpose(cols, (0, 2, 1)).reshape(51, 32)
            tot["transposed"] += wavefronts(idx).sum() + wavefronts(idx + 1).sum()
        tot["ideal"] += 2 * 4 * 51 * NU
        n += 2 * 51 * NU
for k, v in tot.items():
    print("%-10s wavefronts per LDS.128: %.2f" % (k, v / n))
