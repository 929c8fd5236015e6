#!/usr/bin/env python
"""Build kernel variants of liblss_b200.so (same ABI, different -D tuning macros) into build/variants/,
for A/B timing on the GPU box (build/ is git-ignored but travels with gpurun).

    python tools/build_variants.py name1:-DLSS_FWD_M=2 name2:-DLSS_BWD_WIDE=0,-DLSS_BWD_BINS=10 ...
then on the box:  LSS_B200_LIB=build/variants/name1.so python bench.py ...   (tools/run_variants.sh)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lss2_multimodal_nu_b200", "csrc")
OUT = os.path.join(ROOT, "build", "variants")


def main():
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition(":")
        flags = [f for f in flags.split(",") if f]
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
               "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o",
               os.path.join(OUT, name + ".so"), os.path.join(CSRC, "lss_abi.cu")] + flags
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        print(name, "ok" if p.returncode == 0 else "FAILED\n" + out)


if __name__ == "__main__":
    main()
