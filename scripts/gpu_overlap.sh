#!/bin/bash
# N = $1: step time with the overlapped duplicated-node exchange (--opt halo_overlap=1, the default) and without (=0)
N=${1:-2}
mkdir -p gpurun_out
for ov in 1 0; do
  ( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --opt halo_overlap=$ov ) > gpurun_out/overlap_n${N}_$ov.log 2>&1
  echo "== overlap $ov"; grep -o '"value": [0-9.]*' gpurun_out/overlap_n${N}_$ov.log | head -1; grep -o '"step_breakdown_ms[^}]*}' gpurun_out/overlap_n${N}_$ov.log; grep -o '"u_rel_l2": [0-9.e-]*' gpurun_out/overlap_n${N}_$ov.log
done
