#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "column" ) > gpurun_out/pytest_col.log 2>&1
tail -4 gpurun_out/pytest_col.log
timeout 600 python scripts/probe_col.py 256 256 128 6 10 > gpurun_out/probe_col_256.json 2> gpurun_out/probe_col_256.err
cat gpurun_out/probe_col_256.json; tail -3 gpurun_out/probe_col_256.err
timeout 600 python scripts/probe_col.py 128 128 64 5 10 > gpurun_out/probe_col_128.json 2> gpurun_out/probe_col_128.err
tail -1 gpurun_out/probe_col_128.json; tail -3 gpurun_out/probe_col_128.err
