#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an ncu report (source page).

    python tools/ncu_top.py REPORT.ncu-rep KERNEL_REGEX [N]
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
si, ni = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for k, r in enumerate(rows[h + 1:]):
    try:
        n = int(r[ni])
    except Exception:
        continue
    st = sorted(((int(r[i] or 0), c) for i, c in stall_cols), reverse=True)[:2]
    data.append((n, k, r[si].strip(), st))
tot = sum(d[0] for d in data) or 1
print(rows[0][:2], "instructions", len(data), "samples", tot)
agg = {}
for n, k, s, st in data:
    for cnt, c in st:
        agg[c] = agg.get(c, 0) + cnt
print("stall mix:", ", ".join("%s %.0f%%" % (c, 100 * v / tot) for c, v in sorted(agg.items(), key=lambda t: -t[1])[:6]))
for n, k, s, st in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%% @%4d %-70s %s" % (n, 100 * n / tot, k, s[:70], " ".join("%s=%d" % (c[6:], v) for v, c in st if v)))
