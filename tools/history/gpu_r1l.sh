#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_nbody_gpu.py -x -q -m gpu 2>&1 | tail -2
python tools/nbody_bench.py 108 20 f4 | tail -1
python tools/nbody_bench.py 64 20 f8 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 3 -c 1 -o gpurun_out/prof_r1o_nbody_fused python tools/nbody_bench.py 64 3 f4 > gpurun_out/r1o_ncu.log 2>&1
