#!/bin/bash
# end-of-round confirmation: full GPU parity suite, smoke, bench line with the CPU baseline, reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
grep "smoke" gpurun_out/smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/bench.log 2>&1
grep '^{' gpurun_out/bench.log | cut -c1-400
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
grep '^{' gpurun_out/bench_ref.log | cut -c1-200
