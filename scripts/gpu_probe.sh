#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/profile_kernels.py 256 256 128 6 5 > gpurun_out/probe_256.log 2>&1
tail -1 gpurun_out/probe_256.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1
tail -3 gpurun_out/pytest_parity.log
( time timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/bench_1gpu.log 2>&1
tail -4 gpurun_out/bench_1gpu.log | cut -c1-3000
