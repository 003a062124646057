#!/bin/bash
mkdir -p gpurun_out
for o in "mid_lanes=4" "mid_lanes=8" "mid_lanes=16" "quad_nodes=2000000"; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt $o > gpurun_out/bench_opt.log 2>&1
  echo "$o: $(grep -o '"value": [0-9.]*' gpurun_out/bench_opt.log | head -1) $(grep -o '"coarse_levels": [0-9.]*' gpurun_out/bench_opt.log)"
done
