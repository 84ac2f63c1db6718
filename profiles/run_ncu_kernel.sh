#!/bin/bash
# gpurun -- 'bash profiles/run_ncu_kernel.sh TAG REGEX [SKIP] [COUNT]': one `--set full` capture (with source
# counters) of the kernels matching REGEX in a 2-view iteration, after a plain run of the same command exits 0.
TAG=$1; RE=$2; SKIP=${3:-6}; COUNT=${4:-2}
mkdir -p gpurun_out
CMD="python bench.py --views 2 --steps 1 --warmup 3 --lanes 1 --no-e2e --no-cpu --no-render --no-timing"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $COUNT -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
