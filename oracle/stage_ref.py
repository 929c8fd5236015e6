#!/usr/bin/env python
"""Stage the UNMODIFIED reference's Python files into oracle/_ref/ (git-ignored, travels to the GPU box).

    python oracle/stage_ref.py            # needs /root/reference (build container); __graft_entry__.build() calls it

The reference is pure Python: there is nothing to compile, "building" it means copying `src/*.py` and the entry
scripts byte for byte.  oracle/_ref/ is TEST INFRASTRUCTURE: the -m gpu tests and bench.py's `reference_gpu` field
import the reference's own classes from it (through oracle/ref_import.py, with the backbone packages replaced by
oracle/shims/) to run them unpatched and patched on the B200.  Nothing in the product path reads it.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("LSS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
SCRIPTS = ["train.py", "pre_train.py", "predict.py", "pre_train_vovnet.py", "train_vovnet_transformer.py"]


def stage(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "src")):
        if verbose:
            print("stage_ref: %s not present; keeping %s as is" % (SRC, DST))
        return False
    os.makedirs(os.path.join(DST, "src"), exist_ok=True)
    files = [os.path.join("src", f) for f in sorted(os.listdir(os.path.join(SRC, "src"))) if f.endswith(".py")]
    files += [f for f in SCRIPTS if os.path.exists(os.path.join(SRC, f))]
    manifest = []
    for rel in files:
        shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
        with open(os.path.join(DST, rel), "rb") as f:
            manifest.append("%s  %s" % (hashlib.sha256(f.read()).hexdigest(), rel))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print("stage_ref: %d files -> %s" % (len(files), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
