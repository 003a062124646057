#!/bin/bash
mkdir -p gpurun_out
for o in "relax_tab=2" "relax_tab=12" "relax_tab=22" "matvec_tab=14" "matvec_tab=24"; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt $o > gpurun_out/bench_opt.log 2>&1
  echo "$o: $(grep -o '"value": [0-9.]*' gpurun_out/bench_opt.log | head -1) $(grep -o '"relax_fine": [0-9.]*' gpurun_out/bench_opt.log) $(grep -o '"matvec_fine": [0-9.]*' gpurun_out/bench_opt.log) $(grep -o '"coarse_levels": [0-9.]*' gpurun_out/bench_opt.log)"
done
