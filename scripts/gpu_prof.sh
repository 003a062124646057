#!/bin/bash
# ncu evidence for the round: (1) --set full capture of the finest-level smoother pass and matvec (256x256x128),
# (2) launch list (gpu__time_duration) of a bench step at 128x128x64
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ccu_k_relax_tab|ccu_k_matvec_tab' -s 18 -c 3 \
    -o gpurun_out/prof_tab -f python scripts/profile_kernels.py 256 256 128 6 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_tab.ncu-rep --page raw --csv > gpurun_out/prof_tab_raw.csv 2>/dev/null
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --mesh 128x128x64 --levels 5 > gpurun_out/bench_ncu_launches.log 2>&1
wc -l gpurun_out/launches.csv
ls -la gpurun_out/*.ncu-rep
