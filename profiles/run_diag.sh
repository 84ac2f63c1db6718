#!/bin/bash
# gpurun -- 'bash profiles/run_diag.sh TAG "<ENV=..::bench args>" ...': like run_ab.sh without the GPU tests (timing diagnosis
# of deliberately incomplete kernel variants: their results are wrong, only the stage times are read).
TAG=$1; shift
mkdir -p gpurun_out
: > gpurun_out/diag_$TAG.jsonl
for ARGS in "$@"; do
  echo "== $ARGS"
  ENVS=""; case "$ARGS" in *::*) ENVS="${ARGS%%::*}"; ARGS="${ARGS#*::}";; esac
  env $ENVS python bench.py --no-e2e --no-cpu --no-render $ARGS 2> gpurun_out/diag_$TAG.err | tee -a gpurun_out/diag_$TAG.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(round(d['value'], 3), 'iters/s', round(d['ms_per_step'], 2), 'ms', {k: round(v['ms_per_step'], 2) for k, v in d['roofline_stages'].items()})
"
done
