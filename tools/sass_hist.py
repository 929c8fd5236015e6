#!/usr/bin/env python
"""Opcode histogram of one kernel from an ncu report's source page (executed warp instructions and stall
samples per SASS opcode): where a kernel's issue slots go.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<pat> > k.csv ; python tools/sass_hist.py k.csv
"""
import csv
import sys
from collections import Counter


def main(path, top=45):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(r for r in rows if "Instructions Executed" in r)
    ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    body = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[ia].isdigit()]
    tot = sum(int(r[ia]) for r in body)
    tots = sum(int(r[isamp]) for r in body)
    c, s = Counter(), Counter()
    for r in body:
        t = r[isrc].split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op] += int(r[ia]); s[op] += int(r[isamp])
    print("warp instructions %d, SASS lines %d, samples %d" % (tot, len(body), tots))
    for op, n in c.most_common(top):
        print("%-28s %10d %5.1f%%   samples %6d %5.1f%%" % (op, n, 100.0 * n / tot, s[op], 100.0 * s[op] / max(tots, 1)))


if __name__ == "__main__":
    main(sys.argv[1])
