#!/usr/bin/env python3
"""Executed warp-instructions per SOURCE LINE of one kernel: joins the per-SASS-instruction counts of an ncu source
page (`ncu -i X.ncu-rep --page source --csv`) with the line table of the cubin (`nvdisasm -g -c`), by instruction order.

usage: ncu_hot_lines.py SOURCE_PAGE.csv DISASM.txt KERNEL_SUBSTRING WARPS [SRC_DIR]
  DISASM.txt = `cuobjdump -xelf all lib.so; nvdisasm -g -c fb_render.sm_100a.cubin > DISASM.txt` of the SAME build."""
import collections
import csv
import os
import re
import sys


def main():
    page, dis, kernel, warps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    src_dir = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fuzzyblue_b200", "csrc")
    lines = open(dis).read().split("\n")
    start = next(i for i, ln in enumerate(lines) if ln.startswith(".text.") and kernel in ln)
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.")), len(lines))
    cur, instrs = None, []
    for ln in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            instrs.append(cur)
    rows = list(csv.reader(open(page)))
    blocks, hdr = [], None
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            blocks.append([])
        elif r and r[0].startswith("0x") and blocks:
            blocks[-1].append(r)
    data = next(b for b in blocks if len(b) == len(instrs))
    iE = hdr.index("Instructions Executed")
    by = collections.Counter()
    for c, r in zip(instrs, data):
        by[c] += int(r[iE])
    tot = sum(by.values())
    print(f"# {kernel}: {len(instrs)} SASS instructions, {tot / warps:.1f} executed per warp ({warps:.0f} warps)")
    cache = {}
    for (k, v) in by.most_common(50):
        text = ""
        if k:
            if k[0] not in cache:
                p = os.path.join(src_dir, k[0])
                cache[k[0]] = open(p).read().split("\n") if os.path.exists(p) else []
            if 0 < k[1] <= len(cache[k[0]]):
                text = cache[k[0]][k[1] - 1].strip()[:100]
        print(f"{v / warps:8.1f}  {k[0] if k else '?'}:{k[1] if k else 0:<5d} {text}")


if __name__ == "__main__":
    main()
