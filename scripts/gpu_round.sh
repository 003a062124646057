#!/bin/bash
# full GPU test suite + probe + bench variants
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/probe_col.py 256 256 128 6 10 > gpurun_out/probe_col_256.json 2> gpurun_out/probe_col_256.err
tail -1 gpurun_out/probe_col_256.json | cut -c1-600
for v in "relax_col=1" "relax_col=0"; do
  timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --opt $v > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  tail -c 3000 gpurun_out/bench_$v.json; tail -3 gpurun_out/bench_$v.err
done
