#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_dropin.py -q -m gpu -rs -k "regional_sphere" ) > gpurun_out/pytest_rs.log 2>&1
tail -30 gpurun_out/pytest_rs.log
