#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash profiles/run_profile.sh TAG [full]'): GPU tests, the default bench,
# the ncu launch list of a 4-view iteration and one `--set full` capture of a 2-view iteration.
# Outputs land in gpurun_out/; summaries are made afterwards with profiles/summarize_ncu.py.
TAG=${1:-x}
MODE=${2:-all}
mkdir -p gpurun_out
if [ "$MODE" = all ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
  tail -2 gpurun_out/pytest_gpu_$TAG.log
  ( time python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err ) 2>&1 | tail -3
  ( time python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err ) 2>&1 | tail -3
fi
SMALL="--views 4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-render"
python bench.py $SMALL > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(adam|blend|finalize|fit_loss|gacc|gbuf|preprocess|cs_|scan_|radix|emit|ranges|units)' -c 800 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py $SMALL --no-timing > gpurun_out/ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(adam|blend|finalize|fit_loss|gacc|gbuf|preprocess|cs_|scan_|radix|emit|ranges|units)' -s ${SKIP:-70} -c ${COUNT:-40} -o gpurun_out/prof_$TAG -f \
    python bench.py --views 2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-render --no-timing > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -8
