"""LUT dump tool — the counterpart of /root/reference/examples/dump.rs.

    python -m fuzzyblue_b200.dump [outdir]

Builds the atmosphere of dump.rs (:95-107: default Earth, scattering 16x64x16x4, 4 orders), reads the three tables
back in the linear layout of :175-193 and writes `transmittance.exr`, `irradiance.exr`, `scattering.exr` with the
channel naming of :331-367 (`R,G,B,A` for 2-D tables, `{layer}.{R,G,B,A}` per r layer for the 3-D one; FLOAT for the
RGBA32F tables, HALF for the RGBA16F one) plus raw `.npy` copies.  The EXR writer is a minimal scan-line,
uncompressed OpenEXR 2 writer (no dependency)."""
from __future__ import annotations

import os
import struct
import sys
from typing import Dict

import numpy as np

HALF, FLOAT = 1, 2


def _attr(name: str, typ: str, payload: bytes) -> bytes:
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def write_exr(path: str, channels: Dict[str, np.ndarray]) -> None:
    """`channels`: name -> [height][width] array of float16 (stored HALF) or float32 (stored FLOAT)."""
    names = sorted(channels)                      # the channel list must be sorted by name
    h, w = channels[names[0]].shape
    chlist = b""
    for n in names:
        a = channels[n]
        assert a.shape == (h, w) and a.dtype in (np.float16, np.float32), (n, a.shape, a.dtype)
        chlist += n.encode() + b"\0" + struct.pack("<iB3xii", HALF if a.dtype == np.float16 else FLOAT, 0, 1, 1)
    chlist += b"\0"
    box = struct.pack("<4i", 0, 0, w - 1, h - 1)
    header = (struct.pack("<ii", 20000630, 2) + _attr("channels", "chlist", chlist) + _attr("compression", "compression", b"\0")
              + _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box)
              + _attr("lineOrder", "lineOrder", b"\0") + _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
              + _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
              + _attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    row_bytes = sum(channels[n].dtype.itemsize for n in names) * w
    offset0 = len(header) + 8 * h
    with open(path, "wb") as f:
        f.write(header)
        f.write(struct.pack(f"<{h}Q", *[offset0 + y * (8 + row_bytes) for y in range(h)]))
        for y in range(h):
            f.write(struct.pack("<ii", y, row_bytes))
            for n in names:
                f.write(np.ascontiguousarray(channels[n][y]).astype(channels[n].dtype.newbyteorder("<"), copy=False).tobytes())


def table_channels(table: np.ndarray) -> Dict[str, np.ndarray]:
    """[h][w][4] -> R,G,B,A;  [d][h][w][4] -> '{layer}.R' ... (dump.rs channel_name)."""
    if table.ndim == 3:
        return {c: table[..., i] for i, c in enumerate("RGBA")}
    return {f"{z}.{c}": table[z, ..., i] for z in range(table.shape[0]) for i, c in enumerate("RGBA")}


def dump(outdir: str, device: int = 0, **dims) -> Dict[str, np.ndarray]:
    from . import api
    params = api.Parameters(**(dims or dict(scattering_r_size=16, scattering_mu_size=64, scattering_mu_s_size=16, scattering_nu_size=4)))
    T, S, E = api.precompute_host(api.Builder(device), params)
    os.makedirs(outdir, exist_ok=True)
    out = {"transmittance": T, "irradiance": E, "scattering": S}
    for name, t in out.items():
        write_exr(os.path.join(outdir, name + ".exr"), table_channels(t))
        np.save(os.path.join(outdir, name + ".npy"), t)
    return out


if __name__ == "__main__":
    dump(sys.argv[1] if len(sys.argv) > 1 else ".")
    print("wrote transmittance / irradiance / scattering (.exr, .npy)")
