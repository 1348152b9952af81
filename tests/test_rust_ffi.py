"""The Rust shim (rust/src/lib.rs) cannot be compiled in this image (no cargo / rustc), so its FFI block is
machine-checked here instead: every `extern "C"` declaration and every `#[repr(C)]` struct of the shim is parsed and
compared with include/fuzzyblue.h (names, arity, argument and return types) and with the ctypes mirror (struct sizes);
the public API is checked against the reference's signatures (/root/reference/src/precompute.rs:61-68, :1077-1081,
:2147-2211, src/render.rs:34-40, :194, :209-215) by name, `unsafe`-ness and argument count."""
import ctypes
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


RUST = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()


def rust_externs():
    block = RUST[RUST.index('extern "C" {'):]
    block = block[:block.index("\n    }\n")]
    out = {}
    for m in re.finditer(r"pub fn (fb_[a-z0-9_]+)\((.*?)\)(?:\s*->\s*([^;]+))?;", block):
        args = [a.split(":", 1)[1].strip() for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = (args, (m.group(3) or "").strip())
    return out


def test_extern_block_matches_the_header():
    g = _gen()
    want = {}
    for name, ret, params in g.c_prototypes():
        want[name] = ([g.rust_type(t) for t, _ in params], "" if ret == "void" else g.rust_type(ret))
    got = rust_externs()
    assert sorted(got) == sorted(want), (sorted(set(want) - set(got)), sorted(set(got) - set(want)))
    assert len(got) >= 60
    for name in want:
        assert got[name] == want[name], (name, got[name], want[name])


PRIM = {"f32": 4, "i32": 4, "u32": 4, "usize": 8, "u64": 8, "f64": 8}


def rust_struct_sizes():
    """size / alignment of every #[repr(C)] struct of the ffi module under the C layout rules."""
    sizes = {}
    ffi = RUST[RUST.index("pub mod ffi {"):RUST.index('extern "C" {')]
    for m in re.finditer(r"#\[repr\(C\)\].*?pub struct (\w+)\s*\{(.*?)\}", ffi, flags=re.S):
        name, body = m.group(1), re.sub(r"//[^\n]*", "", m.group(2))
        off, align = 0, 1
        for f in re.finditer(r"pub \w+:\s*([^,]+?)\s*(?:,|$)", body.strip() + ","):
            t = f.group(1).strip()
            def sz(t):
                mm = re.match(r"\[(.+);\s*(\d+)\]$", t)
                if mm:
                    s, a = sz(mm.group(1).strip())
                    return s * int(mm.group(2)), a
                if t in PRIM:
                    return PRIM[t], PRIM[t]
                return sizes[t]
            s, a = sz(t)
            off = (off + a - 1) // a * a + s
            align = max(align, a)
        sizes[name] = ((off + align - 1) // align * align, align)
    return sizes


def test_repr_c_structs_have_the_c_sizes():
    from fuzzyblue_b200 import api, sharded
    s = rust_struct_sizes()
    want = {"FbDensityProfileLayer": api.FbDensityProfileLayer, "FbDensityProfile": api.FbDensityProfile, "FbParams": api.FbParams,
            "FbDrawParams": api.FbDrawParams, "FbExtent2D": api.FbExtent2D, "FbExtent3D": api.FbExtent3D,
            "FbExportLayout": api.FbExportLayout, "FbShardStep": sharded.FbShardStep}
    assert sorted(s) == sorted(want)
    for name, ct in want.items():
        assert s[name][0] == ctypes.sizeof(ct), (name, s[name], ctypes.sizeof(ct))
    assert s["FbParams"][0] == 320 and s["FbDrawParams"][0] == 92          # precompute.rs:1675, render.rs:231


def _methods(type_name):
    """{method: (is_unsafe, n_args without self)} of `impl ... Type { ... }` blocks in the shim."""
    out = {}
    for m in re.finditer(r"\nimpl (?:<[^>]*>\s*)?%s \{(.*?)\n\}" % type_name, RUST, flags=re.S):
        for f in re.finditer(r"pub (unsafe )?fn (\w+)(?:<[^>]*>)?\((.*?)\)\s*(?:->|\{)", m.group(1), flags=re.S):
            args = [a for a in re.split(r",(?![^<]*>)", f.group(3)) if a.strip() and "self" not in a.split(":")[0]]
            out[f.group(2)] = (bool(f.group(1)), len(args))
    return out


def test_public_api_has_the_reference_signatures():
    # the six re-exported names, src/lib.rs:8-12
    for name in ("Atmosphere", "Builder", "Parameters", "PendingAtmosphere", "DrawParameters", "Renderer"):
        assert re.search(r"pub struct %s\b" % name, RUST), name
    b, a, p, r, par = _methods("Builder"), _methods("Atmosphere"), _methods("PendingAtmosphere"), _methods("Renderer"), _methods("Parameters")
    assert b["new"] == (False, 6)                               # precompute.rs:61-68
    assert a["build"] == (True, 3)                              # :1077-1081 (unsafe, builder + cmd + &params)
    assert p["acquire_ownership"] == (True, 3)                  # :2147-2152
    assert p["atmosphere"] == (True, 0) and p["assert_ready"] == (True, 0)    # :2203-2211
    for t in ("transmittance", "scattering", "irradiance"):    # :2075-2101
        assert a[t] == (False, 0) and a[t + "_view"] == (False, 0) and a[t + "_extent"] == (False, 0)
        assert par[t + "_extent"] == (False, 0)                 # :771-793
    assert r["new"] == (False, 5)                               # render.rs:34-40
    assert r["set_depth_buffer"] == (True, 2)                   # :194
    assert r["draw"] == (False, 4)                              # :209-215
    # the misspelled fields of the reference are part of its API (precompute.rs:756, :761)
    assert "pub absorbtion_density: DensityProfile" in RUST and "pub absorbtion_extinction: [f32; 3]" in RUST
    # Default = Earth, precompute.rs:849-935
    for lit in ("order: 4", "bottom_radius: 6360.0", "top_radius: 6420.0", "mu_s_min: -0.207912", "0.005802, 0.013558, 0.033100"):
        assert lit in RUST, lit
