#!/bin/bash
# ncu --set full of the column kernels (first matvec launches + first relax launches) at 256x256x128
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^ccu_k_col$' -c 4 \
    -o gpurun_out/prof_col -f python scripts/probe_col.py 256 256 128 6 1 > gpurun_out/ncu_col.log 2>&1
tail -3 gpurun_out/ncu_col.log
ncu -i gpurun_out/prof_col.ncu-rep --page raw --csv > gpurun_out/prof_col_raw.csv 2>/dev/null
ls -la gpurun_out/prof_col*
