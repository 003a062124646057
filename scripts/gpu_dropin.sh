#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_gpu_dropin.py -q -m gpu -rs -k "function_level or marker_bindings or multi_rank" ) > gpurun_out/pytest_dropin.log 2>&1
tail -30 gpurun_out/pytest_dropin.log
