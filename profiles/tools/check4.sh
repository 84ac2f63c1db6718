#!/bin/bash
# gpurun --gpus 4 -- 'bash profiles/tools/check4.sh': the 2-GPU tests (both tails, densify in between) and the driver-style 4-GPU line
python -m pytest tests/test_gpu_fit.py -q -k "two_gpu" 2>&1 | tail -4
bash profiles/run_scale.sh 4 r02f
