#!/bin/bash
# multi-GPU visit: multi-GPU parity tests, then the scaling bench on the GPUs present
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/pytest_multi.log 2>&1
tail -15 gpurun_out/pytest_multi.log
bash scripts/gpu_scale.sh
