"""The EXR writer of the dump tool (examples/dump.rs counterpart): header layout and pixel round trip (CPU), and the
whole tool on the GPU."""
import os
import struct

import numpy as np
import pytest

from fuzzyblue_b200 import dump


def read_exr(path):
    """Minimal reader for what write_exr produces (scan-line, uncompressed)."""
    b = open(path, "rb").read()
    assert struct.unpack_from("<ii", b, 0) == (20000630, 2)
    pos, attrs = 8, {}
    while b[pos] != 0:
        e = b.index(b"\0", pos); name = b[pos:e].decode(); pos = e + 1
        e = b.index(b"\0", pos); typ = b[pos:e].decode(); pos = e + 1
        n, = struct.unpack_from("<i", b, pos); pos += 4
        attrs[name] = (typ, b[pos:pos + n]); pos += n
    pos += 1
    chl, chans, q = attrs["channels"][1], [], 0
    while chl[q] != 0:
        e = chl.index(b"\0", q); nm = chl[q:e].decode(); q = e + 1
        pt, lin, xs, ys = struct.unpack_from("<iB3xii", chl, q); q += 16
        chans.append((nm, pt)); assert (xs, ys) == (1, 1)
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    offs = struct.unpack_from(f"<{h}Q", b, pos)
    out = {nm: np.zeros((h, w), np.float16 if pt == 1 else np.float32) for nm, pt in chans}
    for y in range(h):
        p = offs[y]
        yy, size = struct.unpack_from("<ii", b, p); p += 8
        assert yy == y
        for nm, pt in chans:
            dt = np.dtype("<f2") if pt == 1 else np.dtype("<f4")
            out[nm][y] = np.frombuffer(b, dt, w, p); p += w * dt.itemsize
    return out, attrs, [c[0] for c in chans]


def test_exr_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    t2 = rng.random((5, 7, 4)).astype(np.float32)
    t3 = rng.random((11, 3, 4, 4)).astype(np.float16)
    dump.write_exr(str(tmp_path / "a.exr"), dump.table_channels(t2))
    dump.write_exr(str(tmp_path / "b.exr"), dump.table_channels(t3))
    a, attrs, names = read_exr(tmp_path / "a.exr")
    assert names == ["A", "B", "G", "R"] and attrs["compression"][1] == b"\0"
    for i, c in enumerate("RGBA"):
        assert np.array_equal(a[c], t2[..., i])
    b3, _, names3 = read_exr(tmp_path / "b.exr")
    assert names3 == sorted(names3) and len(names3) == 44 and "10.R" in names3       # dump.rs: "{layer}.{channel}"
    for z in range(11):
        for i, c in enumerate("RGBA"):
            assert np.array_equal(b3[f"{z}.{c}"], t3[z, ..., i])


@pytest.mark.gpu
def test_dump_tool(tmp_path):
    out = dump.dump(str(tmp_path))
    assert out["scattering"].shape == (16, 64, 64, 4) and out["scattering"].dtype == np.float16   # dump.rs:101-107
    for name in ("transmittance", "irradiance", "scattering"):
        assert os.path.getsize(tmp_path / f"{name}.exr") > out[name].nbytes
        assert np.array_equal(np.load(tmp_path / f"{name}.npy"), out[name])
    sc, _, _ = read_exr(tmp_path / "scattering.exr")
    assert np.array_equal(sc["3.G"], out["scattering"][3, ..., 1])
