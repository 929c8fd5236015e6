#!/bin/bash
# ncu evidence for one tag.  usage: tools/prof.sh TAG [KERNEL_REGEX] [SKIP] [COUNT] [CONFIG]
#   1. launch list with per-kernel time / instructions / DRAM and L2 bytes (all kernels of a short run)
#   2. full capture (--set full, source) of the kernels matching KERNEL_REGEX, after SKIP matching launches
TAG=${1:-r2x}; PAT=${2:-"pool_fwd|liftsplat_bwd|plan_|feat_stage"}; SKIP=${3:-42}; CNT=${4:-7}; CFG=${5:-config2}
CMD="python bench.py --config $CFG --no-graph --steps 3 --warmup 3 --repeats 1 --quick --in-flight 1 --sets 2"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/ncu2_$TAG.log
