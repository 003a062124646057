#!/bin/bash
# 8-GPU bench line (2x2x2 subdomains)
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/bench_8gpu.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/bench_8gpu.log | head -1
grep -o '"step_breakdown_ms": {[^}]*}' gpurun_out/bench_8gpu.log
grep -o '"uzawa_iterations": [^]]*]' gpurun_out/bench_8gpu.log
tail -3 gpurun_out/bench_8gpu.log | cut -c1-300
