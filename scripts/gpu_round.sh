#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
( time timeout 1500 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --mesh 128x128x64 --levels 5 > gpurun_out/bench_ncu_launches.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench.log; tail -1 gpurun_out/bench_ref.log; wc -l gpurun_out/launches.csv
