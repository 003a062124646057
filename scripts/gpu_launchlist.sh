#!/bin/bash
# ncu launch list (gpu__time_duration) of the bench command itself at 256x256x128: the first 4000 launches of one step
mkdir -p gpurun_out
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches256.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_ncu_launches256.log 2>&1
wc -l gpurun_out/launches256.csv
