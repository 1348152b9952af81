#!/usr/bin/env python3
"""Rewrites the GLSL syntax of the reference's shader sources that is not C++ (nothing else), so that g++ can compile
them against glsl_compat.h.  Reads /root/reference/shaders/*, writes a build-time scratch directory (removed after the build:
reference sources never enter the repository).

    translate.py SHADER_DIR OUT_DIR

Rewrites, all purely syntactic:
  * `#version` lines and `layout(local_size_*) in;` are dropped;
  * interface blocks `layout(...) uniform Name { decls };` become the declarations themselves, at namespace scope;
    `layout(...) uniform [writeonly] sampler2D|sampler3D|image2D|image3D|subpassInput name;` and
    `layout(...) in|out T name;` become plain (thread-local for the per-invocation ones) variables;
  * parameter qualifiers: `out T x` / `inout T x` -> `T& x`, `in T x` -> `T x`;
  * multi-component swizzles `.rgb .xyz .xy` become accessor calls (single components are struct members);
  * floating-point literals get an `f` suffix (a GLSL `2.0` is a 32-bit float; a C++ `2.0` is a double);
  * `void main()` -> `void shader_main()`.
"""
import os
import re
import sys

TYPES = r"(?:float|int|uint|bool|vec[234]|ivec[23]|uvec[23]|mat4)"
PER_INVOCATION = {"screen_coords", "color_out", "transmittance_out"}


def strip_comments_keep_lines(src):
    """Comments are left alone by every rule below; blank them out for matching, restore afterwards is unnecessary
    because rules only ever ADD characters inside code.  We simply protect comments with placeholders."""
    holders = []

    def keep(m):
        holders.append(m.group(0))
        return "\x00%d\x00" % (len(holders) - 1)

    src = re.sub(r"//[^\n]*|/\*.*?\*/", keep, src, flags=re.S)
    return src, holders


def restore(src, holders):
    return re.sub(r"\x00(\d+)\x00", lambda m: holders[int(m.group(1))], src)


def translate(src):
    src, holders = strip_comments_keep_lines(src)
    src = re.sub(r"^\s*#version[^\n]*\n", "\n", src, flags=re.M)
    src = re.sub(r"layout\s*\(\s*local_size_[^)]*\)\s*in\s*;", "", src)

    def block(m):
        decls = [d.strip() for d in m.group(1).split(";") if d.strip()]
        return "\n".join("static %s;" % d for d in decls)

    src = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+\w+\s*\{([^}]*)\}\s*;", block, src)
    src = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+(?:writeonly\s+|readonly\s+)?(sampler2D|sampler3D|image2D|image3D|subpassInput)\s+(\w+)\s*;",
                 r"static \1 \2;", src)

    def io(m):
        name = m.group(2)
        return ("static thread_local %s %s;" if name in PER_INVOCATION else "static %s %s;") % (m.group(1), name)

    src = re.sub(r"layout\s*\([^)]*\)\s*(?:in|out)\s+(" + TYPES + r")\s+(\w+)\s*;", io, src)
    src = re.sub(r"\b(?:out|inout)\s+(" + TYPES + r")\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"\bin\s+(" + TYPES + r")\s+(\w+)", r"\1 \2", src)
    src = re.sub(r"\.(rgb|xyz|xy)\b", r".\1()", src)
    # float literals: 1.0  .5  2.  1e-3  1.5e3 -> + f (already suffixed ones are left alone)
    src = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])", r"\1f", src)
    src = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", src)
    return restore(src, holders)


def main():
    src_dir, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src_dir)):
        if not name.endswith((".h", ".comp", ".frag")):
            continue
        text = open(os.path.join(src_dir, name)).read()
        with open(os.path.join(out_dir, name), "w") as f:
            f.write("// GENERATED at build time from %s by oracle/glsl_ref/translate.py -- not part of the repository\n" % os.path.join(src_dir, name))
            f.write(translate(text))
        n += 1
    print("translated %d shader files into %s" % (n, out_dir))


if __name__ == "__main__":
    main()
