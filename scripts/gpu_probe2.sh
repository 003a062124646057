#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/profile_kernels.py 256 256 128 6 5 2>&1 | head -1 | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
