"""The one remaining A/B build flag (-DFB_RENDER_IEEE_GUARDS=1: every division / square root of the sky evaluation in
its guarded IEEE form, the cross-check of tools/render_ab.py) must keep compiling for sm_100a: nvcc cross-compiles here
without a GPU.  Objects go to a temporary directory; the product library is not touched.  The variants staged in
round 1 were settled on a B200 in round 2 (profiles/r2_staged_variants_ab.txt): winners adopted, losers deleted."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fuzzyblue_b200", "csrc")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")

CASES = [("fb_render.cu", ["-DFB_RENDER_IEEE_GUARDS=1"])]


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
