python -m pytest tests/test_gpu_tcgen05_oracle.py -q -k crop 2>&1 | tail -2
python bench.py --ext > gpurun_out/r02_ext.json 2> gpurun_out/r02_ext.err; cat gpurun_out/r02_ext.json | cut -c1-900; tail -2 gpurun_out/r02_ext.err
python bench.py --no-cpu --no-render --no-timing --no-e2e --cutoff-sigma 4.5 > gpurun_out/r02_k45.json 2> gpurun_out/r02_k45.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_k45.json').read().strip().splitlines()[-1]); print('k=4.5:', d['value'], d['ms_per_step'], d['pairs'])"
bash profiles/run_scale.sh 2 r02d
