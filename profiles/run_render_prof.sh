#!/bin/bash
# gpurun -- 'bash profiles/run_render_prof.sh TAG': render config (BASELINE configs[2]) alone + its ncu launch list
TAG=$1
mkdir -p gpurun_out
python bench.py --render-only > gpurun_out/render_$TAG.json 2> gpurun_out/render_$TAG.err || { tail -5 gpurun_out/render_$TAG.err; exit 1; }
cat gpurun_out/render_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(blend|finalize|preprocess|cs_|scan_|radix|emit|ranges|units|bin|sort|hist|scatter)' -c 400 --csv \
    --log-file gpurun_out/launches_render_$TAG.csv python bench.py --render-only > gpurun_out/ncu_render_$TAG.log 2>&1
tail -2 gpurun_out/ncu_render_$TAG.log | cut -c1-300
