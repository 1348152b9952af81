#!/usr/bin/env python3
"""Resolve `#if NAME` / `#else` / `#endif` blocks and the `#ifndef NAME / #define NAME v / #endif` default of settled
variant macros in a source file (a minimal unifdef: only bare `#if NAME` conditions are touched).

usage: resolve_ifdef.py FILE NAME=0|1 [NAME=0|1 ...]
"""
import re
import sys


def resolve(lines, settled):
    out = []
    stack = []          # per open #if: (kind, keep_now) kind in {"settled", "other"}
    i = 0
    while i < len(lines):
        ln = lines[i]
        s = ln.strip()
        m = re.match(r"#\s*ifndef\s+(\w+)\s*$", s)
        if m and m.group(1) in settled and i + 2 < len(lines) and re.match(r"#\s*define\s+" + m.group(1) + r"\b", lines[i + 1].strip()) \
                and re.match(r"#\s*endif", lines[i + 2].strip()):
            i += 3      # drop the default definition
            continue
        m = re.match(r"#\s*if\s+(\w+)\s*$", s)
        if m and m.group(1) in settled:
            stack.append(["settled", bool(settled[m.group(1)])])
            i += 1
            continue
        if re.match(r"#\s*(if|ifdef|ifndef)\b", s):
            stack.append(["other", None])
        elif re.match(r"#\s*else\b", s) and stack and stack[-1][0] == "settled":
            stack[-1][1] = not stack[-1][1]
            i += 1
            continue
        elif re.match(r"#\s*endif\b", s) and stack:
            kind, _ = stack.pop()
            if kind == "settled":
                i += 1
                continue
        if all(k != "settled" or keep for k, keep in stack):
            out.append(ln)
        i += 1
    return out


def main():
    path = sys.argv[1]
    settled = {}
    for a in sys.argv[2:]:
        n, v = a.split("=")
        settled[n] = int(v)
    with open(path) as f:
        lines = f.readlines()
    out = resolve(lines, settled)
    with open(path, "w") as f:
        f.writelines(out)
    print("%s: %d -> %d lines" % (path, len(lines), len(out)))


if __name__ == "__main__":
    main()
