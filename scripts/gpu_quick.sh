#!/bin/bash
# parity tests that touch the transfers + a short bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_build.py -x -q 2>&1 | tail -4
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_quick.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/bench_quick.log | head -1; grep -o '"step_breakdown_ms[^}]*}' gpurun_out/bench_quick.log; grep -o '"parity[^}]*}' gpurun_out/bench_quick.log
