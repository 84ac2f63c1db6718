#!/bin/bash
# compute-sanitizer over the small end-to-end workload; logs -> gpurun_out/sanitizer_<tool>.log
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python profiles/tools/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_TARGET_OK' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
