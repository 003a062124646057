#!/bin/bash
# round-2 final single-GPU artifacts: ncu --set full of the dominant kernels of the FINAL build (colour-pass smoother, column matvec /
# residual, column smoother), ncu launch list of the bench command, the bench line itself (parity gate + CPU baseline), the reference arm
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^ccu_k_col$|ccu_k_relax_tab' -c 14 \
    -o gpurun_out/prof_final -f python scripts/probe_col.py 256 256 128 6 1 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches256.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/bench_ncu_launches256.log 2>&1
wc -l gpurun_out/launches256.csv
( time timeout 1200 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err
tail -c 1500 gpurun_out/bench_reference_final.json; tail -4 gpurun_out/bench_reference_final.err
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 4000 gpurun_out/bench_final.json; tail -4 gpurun_out/bench_final.err
