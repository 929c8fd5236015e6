#!/bin/bash
# ncu evidence for one tag: launch list + full capture of the hot kernels.  usage: tools/prof.sh TAG [KERNEL_REGEX] [SKIP] [COUNT]
TAG=${1:-r1x}; PAT=${2:-"pool_fwd|liftsplat_bwd|plan_"}; SKIP=${3:-30}; CNT=${4:-6}
CMD="python bench.py --no-graph --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/ncu2_$TAG.log
