#!/bin/bash
# round-2 closing artifacts on one GPU: ncu launch list of the bench command (first 6000 launches), then the bench line itself
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches256.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/bench_ncu_launches256.log 2>&1
wc -l gpurun_out/launches256.csv
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 3000 gpurun_out/bench_final.json; tail -4 gpurun_out/bench_final.err
