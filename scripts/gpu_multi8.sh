#!/bin/bash
# multi-GPU parity (2, 4, 8 subdomains) + bench lines at N = 8 and 4 with the parity gate
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 2400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -rs ) > gpurun_out/pytest_multi8.log 2>&1
tail -15 gpurun_out/pytest_multi8.log
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  tail -c 1500 gpurun_out/bench_n$n.json; tail -2 gpurun_out/bench_n$n.err
done
