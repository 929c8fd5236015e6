#!/usr/bin/env python
"""Basic-block view of one kernel from an ncu source page: runs of consecutive SASS instructions with the
same executed count, with their size, total issue slots and stall samples -- shows which loop the issue
slots of a kernel go to.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<pat> > k.csv ; python tools/sass_blocks.py k.csv [min_share_pct]
"""
import csv
import sys
from collections import Counter


def main(path, min_share=1.0):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(r for r in rows if "Instructions Executed" in r)
    ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    body = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[ia].isdigit()]
    tot = sum(int(r[ia]) for r in body)
    blocks, cur = [], None
    for k, r in enumerate(body):
        n = int(r[ia])
        if cur is None or n != cur["n"]:
            cur = {"n": n, "start": k, "ops": Counter(), "len": 0, "samples": 0}
            blocks.append(cur)
        t = r[isrc].split()
        op = t[1] if t[0].startswith("@") else t[0]
        cur["ops"][op.split(".")[0]] += 1
        cur["len"] += 1
        cur["samples"] += int(r[isamp])
    print("total warp instructions %d in %d SASS lines" % (tot, len(body)))
    for b in blocks:
        share = 100.0 * b["n"] * b["len"] / tot
        if share < min_share:
            continue
        ops = " ".join("%s:%d" % kv for kv in b["ops"].most_common(9))
        print("@%4d len %3d x %8d = %5.1f%%  samples %4d | %s" % (b["start"], b["len"], b["n"], share, b["samples"], ops))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
