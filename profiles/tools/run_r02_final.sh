#!/bin/bash
# gpurun -- 'bash profiles/tools/run_r02_final.sh TAG': final evidence of the round on one B200 -- smoke, GPU suite, default
# bench + reference arm, the unchanged fit script (configs 1-2), ncu launch list + --set full capture of the fit iteration.
TAG=${1:-r02b}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
python __graft_entry__.py smoke 2>&1 | tail -2
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
cp gpurun_out/parity_report.jsonl gpurun_out/parity_report_$TAG.jsonl
( time python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err ) 2>&1 | tail -3
( time python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err ) 2>&1 | tail -3
python bench.py --fit-scripts > gpurun_out/fit_scripts_$TAG.json 2> gpurun_out/fit_scripts_$TAG.err
SMALL="--views 4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-render"
python bench.py $SMALL > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py $SMALL --no-timing > gpurun_out/ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(blend_wsum|gbuf|preprocess_views|preprocess_bwd|cs_scatter|cs_hist|adam)' -s ${SKIP:-30} -c ${COUNT:-14} -o gpurun_out/prof_$TAG -f \
    python bench.py --views 2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-render --no-timing > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
