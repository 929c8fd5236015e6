"""Run an unmodified reference script with the B200 hot path installed.

    python -m lss2_multimodal_nu_b200.run train.py --dataroot ... --bsize 8
    python -m lss2_multimodal_nu_b200.run predict.py ...
    LSS_STATIC_CALIB=1 python -m lss2_multimodal_nu_b200.run predict.py ...   # fixed rig: build the plan once

The script's directory is put on sys.path (so ``from src... import`` resolves as
it does when the script is run directly), the reference classes are patched
(patch.install_reference_classes), a global forward pre-hook patches any other
module with the lift-splat surface at its first call (patch.install_everywhere:
``PreTrainingModel`` is defined inside pre_train_vovnet.py) and the script is
executed with runpy as ``__main__``.  The reference files are not touched.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    sys.path.insert(0, os.path.dirname(script))
    from . import patch
    n = patch.install_reference_classes()
    # ... and everything with the lift-splat surface at its first forward call: the model class of
    # pre_train_vovnet.py (PreTrainingModel) is defined in the script itself, i.e. in __main__
    patch.install_everywhere()
    if os.environ.get("LSS_STATIC_CALIB"):
        patch.STATIC_BY_DEFAULT = True
    if n == 0:
        print("lss2_multimodal_nu_b200.run: no reference model classes found next to %s; models are patched at "
              "their first forward call" % script, file=sys.stderr)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
