#!/bin/bash
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool racecheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_parity.py -q -x -k "tile_kernels_match and busse_l3 and 0-1 or kernel_variants_agree and busse" > gpurun_out/racecheck.log 2>&1
grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/racecheck.log | tail -3
grep -A6 "hazard" gpurun_out/racecheck.log | head -40
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_parity.py -q -x -k "tile_kernels_match and busse_l3 and 0-1 or kernel_variants_agree and busse" > gpurun_out/synccheck.log 2>&1
grep "ERROR SUMMARY\|passed\|failed" gpurun_out/synccheck.log | tail -3
