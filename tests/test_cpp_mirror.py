"""Compiles tests/cpp_mirror_test.cpp against include/fuzzyblue.hpp and the shared library; runs it on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp_mirror_test")


def compile_it():
    lib = os.path.join(ROOT, "fuzzyblue_b200", "csrc")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp_mirror_test.cpp"),
           "-o", EXE, f"-L{lib}", "-lfuzzyblue_b200", f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.run(cmd, check=True)


def test_cpp_mirror_compiles_and_refuses_without_gpu():
    import torch
    compile_it()
    r = subprocess.run([EXE], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)   # FB_ERR_NO_DEVICE, no fallback


@pytest.mark.gpu
def test_cpp_mirror_smoke_on_gpu():
    compile_it()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp mirror ok" in r.stdout


C_EXE = os.path.join(ROOT, "examples", "bench_precompute")


def compile_c_example():
    lib = os.path.join(ROOT, "fuzzyblue_b200", "csrc")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-O2", "-Wall", "-Werror", f"-I{os.path.join(ROOT, 'include')}", "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "examples", "bench_precompute.c"), "-o", C_EXE, f"-L{lib}", "-lfuzzyblue_b200",
                    f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-lcudart"], check=True)


def test_c_bench_example_compiles_and_refuses_without_gpu():
    """examples/bench_precompute.c (the C counterpart of the reference's benches/precompute.rs) builds against the ABI."""
    import torch
    compile_c_example()
    r = subprocess.run([C_EXE], capture_output=True, text=True)
    assert r.returncode == (0 if torch.cuda.is_available() else 77), (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_bench_example_runs():
    compile_c_example()
    r = subprocess.run([C_EXE], capture_output=True, text=True)
    assert r.returncode == 0 and "ns/iter" in r.stdout, r.stdout + r.stderr
