#!/bin/bash
# L2 working-set experiment: DRAM bytes / L2 hit rate / time of the tile smoother by shape and CTAs per SM
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
for cfg in "2 0" "2 60" "2 130" "3 0" "0 0" "0 100"; do
set -- $cfg
timeout 600 ncu --metrics $M --clock-control none -k regex:'^ccu_k_tile$' -s 25 -c 3 --csv --log-file gpurun_out/ws_$1_$2.csv \
    python scripts/prof_tile.py 256 256 128 6 $1 1 $2 > gpurun_out/ws_$1_$2.log 2>&1
echo "shape $1 pad $2"; python scripts/ws_print.py gpurun_out/ws_$1_$2.csv
done
