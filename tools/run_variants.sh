#!/bin/bash
# A/B timing of the kernel variants in build/variants/ (tools/build_variants.py) on the GPU box:
# one short bench.py run per library, key numbers of each on one line.  usage: tools/run_variants.sh [bench args]
ARGS=${@:-"--steps 200 --warmup 10 --quick"}
for lib in "" build/variants/*.so; do
  name=${lib:-default}
  LSS_B200_LIB=$lib python bench.py $ARGS > gpurun_out/var_$(basename $name .so).json 2> gpurun_out/var_$(basename $name .so).err
  python - "$name" gpurun_out/var_$(basename $name .so).json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().split("\n")[-1])
    k = d["roofline"]["kernels_us"]
    print("%-28s value %7.0f/s (%.1f us)  serial %.1f us  plan %.1f stage %.1f fwd %.1f bwd %.1f" % (
        sys.argv[1], d["value"], d["ms_per_step"] * 1e3, d["serial"]["ms_per_step"] * 1e3, k["plan"], k["stage"], k["fwd"], k["bwd"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
