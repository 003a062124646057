#!/bin/bash
# ncu --set full capture of the finest-level smoother and matvec kernels (256x256x128)
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'ccu_k_relax|ccu_k_matvec' -s 22 -c 4 \
    -o gpurun_out/prof_relax_matvec -f python scripts/profile_kernels.py 256 256 128 6 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_relax_matvec.ncu-rep --page raw --csv > gpurun_out/prof_relax_matvec_raw.csv 2>/dev/null
ls -la gpurun_out/
