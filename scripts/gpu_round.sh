#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants or column" ) > gpurun_out/pytest_col.log 2>&1
tail -3 gpurun_out/pytest_col.log
for v in "relax_full=1" "matvec_full=1" "relax_full=0"; do
  timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity --opt $v > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$v.json"))
    print("$v", "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "relax", round(d["roofline"]["avg_launch_ms"],4), round(d["roofline"]["frac"],3), "matvec", round(d["matvec"]["avg_launch_ms"],4), round(d["matvec"]["frac"],3), {k: round(x,1) for k,x in d["step_breakdown_ms"].items()}, d["uzawa_iterations"])
except Exception as e:
    print("$v failed", e)
PY
  tail -2 gpurun_out/bench_$v.err
done
