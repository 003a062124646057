#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 --no-parity --opt p2p_halo=$v > gpurun_out/bench_p2p${v}_n2.json 2> gpurun_out/bench_p2p${v}_n2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_p2p${v}_n2.json") if l.startswith("{")][-1])
    print("p2p=$v", d["halo_exchange"], "value", round(d["value"],4), {k: round(x,1) for k,x in d["step_breakdown_ms"].items()}, d["roofline"]["launches"])
except Exception as e:
    print("failed", e)
PY
tail -2 gpurun_out/bench_p2p${v}_n2.err | cut -c1-300
done
