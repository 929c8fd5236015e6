#!/usr/bin/env python
"""Print the SASS of one kernel of a built library (no GPU needed), optionally only the innermost loops
(backward branches), with an opcode count per loop.

    python tools/sass_loop.py LIB.so 'pool_fwd_kernel<(bool)1, (int)8, (int)2' [--full]
"""
import re
import subprocess
import sys
from collections import Counter


def kernel_sass(lib, pat):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = subprocess.run(["bash", "-c", "cuobjdump -sass %s | grep 'Function :' | awk '{print $3}' | c++filt" % lib],
                           capture_output=True, text=True).stdout.split("\n")
    chunks = out.split("Function : ")[1:]
    for name, chunk in zip(names, chunks):
        if pat in name:
            return name, chunk
    raise SystemExit("no kernel matches %r; have:\n%s" % (pat, "\n".join(names)))


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    full = "--full" in sys.argv
    name, chunk = kernel_sass(lib, pat)
    ins = []
    for line in chunk.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, "-", len(ins), "instructions")
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?(?:\.ANY)?\s+(?:[!A-Z0-9, ]*?)0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    for s, e in loops:
        ops = Counter()
        for _, t in ins[s:e + 1]:
            tok = t.split()
            op = tok[1] if tok[0].startswith("@") else tok[0]
            ops[op.split(".")[0]] += 1
        print("loop @%d..%d  len %d : %s" % (s, e, e - s + 1, " ".join("%s:%d" % kv for kv in ops.most_common(14))))
    if full:
        for i, (a, t) in enumerate(ins):
            print("%4d %s" % (i, t))


if __name__ == "__main__":
    main()
