#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_gpu_energy.py tests/test_gpu_dropin.py -q -m gpu -rs -k "output_staging" ) > gpurun_out/pytest_obs.log 2>&1
tail -40 gpurun_out/pytest_obs.log
