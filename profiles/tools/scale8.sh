#!/bin/bash
# gpurun --gpus 8 -- 'bash profiles/tools/scale8.sh': C4 at 8 GPUs with the per-phase timeline, fused NVLink-multicast tail vs NCCL tail
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-render --no-e2e --no-timing --timeline --lanes 4 --grad-chunks $3 --comm $2 2> gpurun_out/scale8_$2_$3.err | tail -1 > gpurun_out/scale8_$2_$3.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/scale8_$2_$3.json').read()); print('$2 chunks $3:', round(d['value'],2),'iters/s', round(d['ms_per_step'],3),'ms host', round(d['host_ms_per_step'],2), d['comm'][:30], {k:round(v,3) for k,v in (d['timeline_ms'] or {}).items()})" || tail -5 gpurun_out/scale8_$2_$3.err; }
run 29631 multimem 4
run 29632 multimem 2
run 29633 nccl 2
