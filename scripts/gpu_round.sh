#!/bin/bash
mkdir -p gpurun_out
for v in "l2_persist=1" "l2_persist=0"; do
  for g in "" "--no-graphs"; do
  timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity --opt $v $g > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_tmp.json"))
    print("$v $g", "value", round(d["value"],4), {k: round(x,1) for k,x in d["step_breakdown_ms"].items()}, {k: round(v["ms"],1) for k,v in d.get("level_ms_per_step",{}).items()})
except Exception as e:
    print("$v failed", e)
PY
  tail -2 gpurun_out/bench_tmp.err
  done
done
