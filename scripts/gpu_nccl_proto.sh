#!/bin/bash
# N=2: does the NCCL protocol choice change the per-sweep exchange latency?
mkdir -p gpurun_out
for proto in default LL LL128; do
  if [ $proto = default ]; then unset NCCL_PROTO; else export NCCL_PROTO=$proto; fi
  ( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline ) > gpurun_out/nccl_$proto.log 2>&1
  echo "== $proto"; grep -o '"value": [0-9.]*' gpurun_out/nccl_$proto.log | head -1; grep -o '"step_breakdown_ms[^}]*}' gpurun_out/nccl_$proto.log
done
