#!/bin/bash
# gpurun -- 'bash profiles/run_ab.sh TAG "<bench args A>" "<bench args B>" ...': GPU tests once, then one short
# bench per argument string ("VAR=1 VAR2=x::--flag" sets environment variables for that run) (stage table on stderr is dropped; the JSON lines land in gpurun_out/ab_TAG.jsonl).
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
: > gpurun_out/ab_$TAG.jsonl
for ARGS in "$@"; do
  echo "== $ARGS" 
  ENVS=""; case "$ARGS" in *::*) ENVS="${ARGS%%::*}"; ARGS="${ARGS#*::}";; esac
  env $ENVS python bench.py --no-e2e --no-cpu --no-render $ARGS 2> gpurun_out/ab_$TAG.err | tee -a gpurun_out/ab_$TAG.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(round(d['value'], 3), 'iters/s', round(d['ms_per_step'], 2), 'ms', {k: round(v['ms_per_step'], 2) for k, v in d['roofline_stages'].items()})
"
done
