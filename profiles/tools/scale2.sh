#!/bin/bash
# gpurun --gpus 2 -- 'bash profiles/tools/scale2.sh': the 2-GPU tests of the multi-GPU tail, then the C4 bench with both tails
python -m pytest tests/test_gpu_fit.py -q -k "two_gpu" 2>&1 | tail -4
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-render --no-e2e --no-timing --timeline --comm $2 2> gpurun_out/scale2_$2.err | tail -1 > gpurun_out/scale2_$2.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/scale2_$2.json').read()); print('$2', round(d['value'],2),'iters/s', round(d['ms_per_step'],3),'ms', d['comm'][:40], {k:round(v,3) for k,v in (d['timeline_ms'] or {}).items() if 'chunk' not in k or 'adam' in k})" || tail -5 gpurun_out/scale2_$2.err; }
run 29621 nccl
run 29622 multimem
