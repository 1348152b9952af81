"""The A/B build flags of the sky evaluation (-DFB_RENDER_IEEE_GUARDS=1: every division / square root in its guarded
IEEE form; -DFB_RENDER_SCALAR_BLENDS=1: the packed fp32 pair operations as scalar ones -- the cross-checks of
tools/render_ab.py) must keep compiling for sm_100a: nvcc cross-compiles here without a GPU.  Objects go to a temporary directory; the product library is not touched.  The variants staged in
round 1 were settled on a B200 in round 2 (profiles/r2_staged_variants_ab.txt): winners adopted, losers deleted."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fuzzyblue_b200", "csrc")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")

CASES = [("fb_render.cu", ["-DFB_RENDER_IEEE_GUARDS=1"]), ("fb_render.cu", ["-DFB_RENDER_SCALAR_BLENDS=1"])]


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not installed")
@pytest.mark.parametrize("src,flags", CASES, ids=lambda v: v if isinstance(v, str) else "+".join(f[2:] for f in v))
def test_staged_variant_compiles(tmp_path, src, flags):
    out = tmp_path / (src + ".o")
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOSTCXX, "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off", "--expt-relaxed-constexpr", *flags,
           "-c", os.path.join(CSRC, src), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.stat().st_size > 0


@pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"), reason="cuobjdump not installed")
def test_product_kernels_use_packed_fp32_pairs():
    """the density kernels and the sky evaluation issue their channel-parallel arithmetic as packed fp32
    pairs (FFMA2, sm_100 only).  Checked in the SASS of the built library; the row-shared kernels, where the pairing
    measured slower, must not carry it."""
    lib = os.path.join(CSRC, "libfuzzyblue_b200.so")
    if not os.path.exists(lib):
        pytest.skip("library not built")
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([tool, "-sass", lib], capture_output=True, text=True, timeout=600).stdout
    counts, name = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            counts[name] = 0
        elif name and "FFMA2" in line:
            counts[name] += 1
    dens = {k: v for k, v in counts.items() if "k_density_main" in k}
    sky = {k: v for k, v in counts.items() if "k_render_sky" in k and "Lb1ELb" in k.split("k_render_sky")[1][:24]}
    rows = {k: v for k, v in counts.items() if "k_multiple_scattering" in k or "k_single_scattering" in k}
    assert len(dens) == 4 and all(v >= 100 for v in dens.values()), dens
    assert sky and any(v >= 40 for v in sky.values()), sky
    assert rows and all(v == 0 for v in rows.values()), rows
