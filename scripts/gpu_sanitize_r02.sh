#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels: column kernels (parity tests), Rsphere / imposed-velocity operator construction,
# observables; racecheck on the column kernels' shared-memory ring
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_build.py -q -x -k "regional_sphere and off or imposed_velocity or rheologies" > gpurun_out/sanitize_build.log 2>&1
grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize_build.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_parity.py -q -x -k "column_kernels_match and busse_l3 or transfers_all_levels and busse_l3" > gpurun_out/sanitize_col.log 2>&1
grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize_col.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_energy.py -q -x -k "stress or observables or output" > gpurun_out/sanitize_obs.log 2>&1
grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize_obs.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_parity.py -q -x -k "column_kernels_match and busse_l3 or transfers_all_levels and busse_l3" > gpurun_out/racecheck_col.log 2>&1
grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/racecheck_col.log | tail -3
