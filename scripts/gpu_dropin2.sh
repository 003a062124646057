#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_dropin.py -q -m gpu -rs -k "multi_rank" ) > gpurun_out/pytest_dropin_multi.log 2>&1
tail -30 gpurun_out/pytest_dropin_multi.log
