#!/bin/bash
mkdir -p gpurun_out
( time timeout 3000 python -m pytest tests -q -m gpu -rs ) > gpurun_out/pytest_gpu.log 2>&1
tail -25 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
