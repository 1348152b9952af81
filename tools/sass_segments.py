"""Segments of an `ncu --page source --csv` dump (SASS view) by execution count: which address ranges carry the
instructions and the stall samples.  usage: tools/sass_segments.py FILE.csv [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isrc, ie, iss = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[ia], 16), r[isrc].strip(), int(r[ie]), int(r[iss])))
    except (ValueError, IndexError):
        pass
base = data[0][0]
tot, tots = sum(d[2] for d in data), sum(d[3] for d in data)
print("total warp-inst", tot, "stall samples", tots, "sass lines", len(data))
seg, cur = [], None
for a, s, e, sm in data:
    if cur and abs(e - cur["e"]) <= 0.15 * max(e, cur["e"], 1):
        cur["n"] += 1; cur["inst"] += e; cur["samp"] += sm; cur["end"] = a - base
        op = s.split()[0] if not s.startswith("@") else s.split()[1]
        cur["ops"][op.split(".")[0]] = cur["ops"].get(op.split(".")[0], 0) + 1
    else:
        cur = {"start": a - base, "end": a - base, "e": e, "n": 1, "inst": e, "samp": sm, "ops": {}}
        seg.append(cur)
for s in seg:
    if s["inst"] > thr * tot or s["samp"] > thr * tots:
        ops = sorted(s["ops"].items(), key=lambda kv: -kv[1])[:8]
        print(f"{s['start']:#7x}-{s['end']:#7x} n={s['n']:4d} exec={s['e']:9d} inst={s['inst'] / tot * 100:5.1f}% "
              f"stall={s['samp'] / tots * 100:5.1f}%  " + " ".join(f"{k}:{v}" for k, v in ops))
