#!/bin/bash
# column-resident kernels on the GPU: parity (small meshes, every level), properties + timings at bench size
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "column" ) > gpurun_out/pytest_col.log 2>&1
tail -15 gpurun_out/pytest_col.log
timeout 300 python scripts/probe_col.py 64 64 32 4 5 > gpurun_out/probe_col_64.json 2> gpurun_out/probe_col_64.err
cat gpurun_out/probe_col_64.json; tail -3 gpurun_out/probe_col_64.err
timeout 600 python scripts/probe_col.py 256 256 128 6 10 > gpurun_out/probe_col_256.json 2> gpurun_out/probe_col_256.err
cat gpurun_out/probe_col_256.json; tail -3 gpurun_out/probe_col_256.err
