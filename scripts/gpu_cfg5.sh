#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5_n1.err
tail -c 2500 gpurun_out/bench_cfg5_n1.json; tail -5 gpurun_out/bench_cfg5_n1.err
