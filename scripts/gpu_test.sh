#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/bench_1gpu.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/bench_1gpu.log | head -1; grep -o '"step_breakdown_ms": {[^}]*}' gpurun_out/bench_1gpu.log
