#!/bin/bash
# 8-GPU session: parity at 2/4/8 subdomains, multi-rank drop-in, bench at N=8 (peer-memory vs NCCL halo), N=4, config 5 at N=8
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py -q -m gpu -rs -k "subdomains or markers_change or multi_rank" ) > gpurun_out/pytest_multi8.log 2>&1
tail -12 gpurun_out/pytest_multi8.log
run() { # n tag opts...
  n=$1; tag=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 3 --warmup 2 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", d.get("halo_exchange"), "value", round(d["value"],4) if d["value"]<1e3 else d["value"], {k: round(x,1) for k,x in d.get("step_breakdown_ms",{}).items()}, d.get("parity"), d.get("checks"))
except Exception as e:
    print("$tag failed", e)
PY
  tail -2 gpurun_out/bench_$tag.err | cut -c1-300
}
run 8 n8_p2p1 --opt p2p_halo=1
run 8 n8_p2p0 --opt p2p_halo=0 --no-parity
run 4 n4_p2p1 --opt p2p_halo=1
run 4 n4_p2p0 --opt p2p_halo=0 --no-parity
run 8 cfg5_n8 --config 5
