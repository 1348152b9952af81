"""Static instruction mix of one kernel in an object file: tools/sass_count.py OBJ SUBSTRING
Counts SASS instructions per opcode class (the sky kernel is one straight-line body with a few branches, so the static
count of the hot variant is a usable proxy for issue slots per pixel before spending GPU time)."""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
name, counts = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); counts[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        counts[name][m.group(1).split(".")[0]] += 1
for n, c in counts.items():
    if pat in n:
        tot = sum(c.values())
        xu = sum(c[k] for k in ("MUFU", "FRND", "F2I", "I2F", "F2F", "I2FP", "F2IP"))
        print(n[:110]); print("  total", tot, " XU-class", xu, " ", dict(c.most_common(14)))
