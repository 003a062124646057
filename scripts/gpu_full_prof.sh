#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/probe_full.py 256 256 128 6 10 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^ccu_k_relax_(full|tab)$' -s 40 -c 2 \
    -o gpurun_out/prof_full -f python scripts/probe_full.py 256 256 128 6 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^ccu_k_relax_full$' -s 8 -c 1 \
    -o gpurun_out/prof_full2 -f python scripts/probe_full.py 256 256 128 6 1 >> gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_full.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_full2.ncu-rep --page raw --csv > gpurun_out/prof_full2_raw.csv 2>/dev/null
ls -la gpurun_out/prof_full*
