#!/bin/bash
# gpurun -- 'bash profiles/tools/run_r02.sh TAG': the round-2 evidence in one call -- GPU tests, default bench, reference arm,
# ncu launch lists (fit iteration + render config) and --set full captures of both.  Outputs -> gpurun_out/.
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
( time python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err ) 2>&1 | tail -3
( time python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err ) 2>&1 | tail -3
K='^(adam|blend|finalize|fit_loss|gacc|gbuf|preprocess|cs_|scan_|radix|emit|ranges|units|seg|slab|sorted|group|depth|u8)'
SMALL="--views 4 --steps 1 --warmup 3 --no-e2e --no-cpu --no-render"
python bench.py $SMALL > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py $SMALL --no-timing > gpurun_out/ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(blend_wsum|gbuf|preprocess_views|preprocess_bwd|cs_scatter|cs_hist)' -s ${SKIP:-30} -c ${COUNT:-12} -o gpurun_out/prof_$TAG -f \
    python bench.py --views 2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-render --no-timing > gpurun_out/ncu_full_$TAG.log 2>&1
python bench.py --render-only > gpurun_out/render_$TAG.json 2> gpurun_out/render_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_render_$TAG.csv python bench.py --render-only > gpurun_out/ncu_render_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(blend_sorted|group_sort|cs_|slab|preprocess_kernel|finalize_sorted)' -s ${RSKIP:-24} -c ${RCOUNT:-13} -o gpurun_out/prof_render_$TAG -f \
    python bench.py --render-only > gpurun_out/ncu_render_full_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
