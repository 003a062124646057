#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graphs ) > gpurun_out/bench_levels.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/bench_levels.log | head -1; grep -o '"step_breakdown_ms.*' gpurun_out/bench_levels.log | cut -c1-1600
