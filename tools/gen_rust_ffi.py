#!/usr/bin/env python3
"""Print the Rust `extern "C"` declarations for every FB_API function of include/fuzzyblue.h (used to write the `ffi`
module of rust/src/lib.rs; tests/test_rust_ffi.py checks the two stay in step)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCALARS = {"int": "c_int", "uint32_t": "u32", "uint64_t": "u64", "size_t": "usize", "float": "f32", "double": "f64",
           "char": "c_char", "void": "c_void", "int32_t": "i32"}


def c_prototypes(text=None):
    """[(name, return C type, [(C type, name)])] for every FB_API declaration."""
    text = text or open(os.path.join(ROOT, "include", "fuzzyblue.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out = []
    for m in re.finditer(r"FB_API\s+([^;(]*?)\b(fb_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)
                params.append((mm.group(1).strip(), mm.group(2)))
        out.append((name, ret, params))
    return out


def rust_type(c: str) -> str:
    """`const FbParams*` -> `*const FbParams`, `FbPending**` -> `*mut *mut FbPending`, `const void**` -> `*mut *const c_void`."""
    c = c.replace(" *", "*").strip()
    stars = len(c) - len(c.rstrip("*"))
    base = c.rstrip("*").strip()
    const = base.startswith("const ")
    base = base[6:].strip() if const else base
    t = SCALARS.get(base, base)
    for i in range(stars):
        inner_const = const and i == 0
        t = ("*const " if inner_const else "*mut ") + t
    return t


def rust_decl(name, ret, params) -> str:
    args = ", ".join(f"{'r#' + n if n in ('type', 'ref', 'in') else n}: {rust_type(t)}" for t, n in params)
    r = "" if ret == "void" else f" -> {rust_type(ret)}"
    return f"pub fn {name}({args}){r};"


if __name__ == "__main__":
    for p in c_prototypes():
        sys.stdout.write("        " + rust_decl(*p) + "\n")
