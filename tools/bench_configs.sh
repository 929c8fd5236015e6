#!/bin/bash
# The other BASELINE.json shapes (lift-splat stage alone): config 4 (B=16, 256x704, D=59, C=80, 200x200) and
# config 5 (B=32, D=118, C=128, 512x512); one JSON line each into gpurun_out/, summary on stdout.
# usage: tools/bench_configs.sh [TAG]    (under torchrun for N > 1: tools/bench_configs.sh TAG N)
TAG=${1:-r2}; N=${2:-1}
for c in config4 config5; do
  out=gpurun_out/${TAG}_${c}_n${N}.json
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --config $c --steps 20 --warmup 3 --sets 2 --in-flight 2 --repeats 7 --no-cpu-baseline --no-reference-gpu --e2e-steps 20 > $out 2> ${out%.json}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $c --steps 20 --warmup 3 --sets 2 --in-flight 2 --repeats 7 --no-cpu-baseline --no-reference-gpu --e2e-steps 20 > $out 2> ${out%.json}.err
  fi
  python - $out <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    r = d["roofline"]
    print("%s n=%d: %7.0f samples/s  %.1f us/step (serial %.1f, cached plan %.1f)  kernels %s  fwd frac %.3f bwd frac %.3f step frac %.3f" % (
        d["config"]["workload"][:60], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["serial"]["ms_per_step"] * 1e3,
        d["cached_plan"]["ms_per_step"] * 1e3, r["kernels_us"], r["frac"], d["roofline_bwd"]["frac"], r["step_frac_of_hbm_roofline"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
