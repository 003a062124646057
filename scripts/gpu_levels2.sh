#!/bin/bash
# per-level times (graphs off) and the graph-on step time at N=1 and N=2
mkdir -p gpurun_out
( timeout 500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graphs ) > gpurun_out/levels_n1.log 2>&1
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --no-graphs ) > gpurun_out/levels_n2.log 2>&1
for f in gpurun_out/levels_n1.log gpurun_out/levels_n2.log; do grep -o '"value": [0-9.]*' $f | head -1; grep -o '"step_breakdown_ms.*' $f | cut -c1-1400; done
