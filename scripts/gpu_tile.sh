#!/bin/bash
# first GPU visit of the tile-resident kernels: parity, timings by shape/hint, one bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tile" ) > gpurun_out/pytest_tile.log 2>&1
tail -5 gpurun_out/pytest_tile.log
timeout 600 python scripts/probe_tile.py 256 256 128 6 5 > gpurun_out/probe_tile.json 2> gpurun_out/probe_tile.err
cat gpurun_out/probe_tile.json | cut -c1-3000; tail -3 gpurun_out/probe_tile.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_tile.log 2>&1
tail -1 gpurun_out/bench_tile.log | cut -c1-2500
