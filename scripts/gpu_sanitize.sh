#!/bin/bash
# compute-sanitizer memcheck over the kernels added this round (small meshes)
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "tile_kernels_match and busse_l3 and 0-1 or kernel_variants_agree and busse or conj_grad" > gpurun_out/sanitize_parity.log 2>&1
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_parity.log; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize_parity.log | tail -3
grep -B2 -A12 "Invalid" gpurun_out/sanitize_parity.log | head -60
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 20 python -m pytest tests/test_gpu_markers.py tests/test_gpu_energy.py -q -x -k "change_subdomain or phase or heating or heat_flux" > gpurun_out/sanitize_energy.log 2>&1
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_energy.log; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitize_energy.log | tail -3
grep -B2 -A12 "Invalid" gpurun_out/sanitize_energy.log | head -60
