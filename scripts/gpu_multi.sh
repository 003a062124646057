#!/bin/bash
# 2-GPU visit: single-GPU regression tests, multi-GPU parity test, 2-GPU bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu2.log 2>&1
tail -15 gpurun_out/pytest_gpu2.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/bench_2gpu.log 2>&1
tail -4 gpurun_out/bench_2gpu.log
( time timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/bench_1gpu.log 2>&1
tail -2 gpurun_out/bench_1gpu.log
