#!/bin/bash
# gpurun --gpus N -- 'bash profiles/run_scale.sh N TAG': the bench exactly as the driver launches it at N GPUs.
N=$1; TAG=$2
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
fi
echo "rc=$?"; tail -c 600 gpurun_out/scale_${TAG}_n$N.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/scale_${TAG}_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],2), 'iters/s', round(d['ms_per_step'],2),'ms', 'e2e', d['e2e'] and round(d['e2e']['value'],2))
"
