#!/bin/bash
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-render --no-e2e --no-timing --timeline --lanes $2 --grad-chunks $3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lanes',$2,'chunks',$3,'iters/s',round(d['value'],2),'ms',round(d['ms_per_step'],3),'host_ms',round(d['host_ms_per_step'],3), {k:round(v,3) for k,v in (d['timeline_ms'] or {}).items()})"; }
run 29602 4 4
run 29603 4 2
