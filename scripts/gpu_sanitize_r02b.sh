#!/bin/bash
# memcheck over the regional-spherical kernels added late in round 2 (energy, heating, observables, stress, phase change, K.VB)
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_energy.py tests/test_gpu_build.py -q -x -k "regional_sphere or (stress_and and Rsphere) or (imposed_velocity and Rsphere)" > gpurun_out/sanitize_rsphere.log 2>&1
grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize_rsphere.log | tail -3
