#!/bin/bash
# how long does the unmodified reference take on this box's host cores at the larger meshes?
mkdir -p gpurun_out
nproc; free -g | head -2
( time timeout 900 python bench.py --impl reference --ref-mesh 128x128x64 --ref-levels 5 --steps 1 --warmup 0 ) > gpurun_out/ref_128.log 2>&1
tail -5 gpurun_out/ref_128.log
( time timeout 1500 python bench.py --impl reference --ref-mesh 256x256x128 --ref-levels 6 --steps 1 --warmup 0 ) > gpurun_out/ref_256.log 2>&1
tail -5 gpurun_out/ref_256.log
