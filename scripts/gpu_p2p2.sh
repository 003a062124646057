#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -rs -k "subdomains" ) > gpurun_out/pytest_multi2.log 2>&1
tail -8 gpurun_out/pytest_multi2.log
