#!/bin/bash
# scaling bench on the GPUs of this box: N = 2, 4 (and 8 when present)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 2 4 8; do
  if [ "$N" -le "$NG" ]; then
    ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
        bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/bench_${N}gpu.log 2>&1
    grep -o '"value": [0-9.]*' gpurun_out/bench_${N}gpu.log | head -1
    grep -o '"step_breakdown_ms": {[^}]*}' gpurun_out/bench_${N}gpu.log
    grep -o '"uzawa_iterations": [^]]*]' gpurun_out/bench_${N}gpu.log
  fi
done
