#!/bin/bash
mkdir -p gpurun_out
NBODY_KERNEL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 3 -c 1 -o gpurun_out/prof_r1i_nbody_v1 python tools/nbody_bench.py 64 3 f4 > gpurun_out/r1i_ncu1.log 2>&1
NBODY_KERNEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_list -s 3 -c 1 -o gpurun_out/prof_r1i_nbody_v2 python tools/nbody_bench.py 64 3 f4 > gpurun_out/r1i_ncu2.log 2>&1
NBODY_KERNEL=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 12 --csv --log-file gpurun_out/r1i_nbody_launches.csv python tools/nbody_bench.py 108 3 f4 > gpurun_out/r1i_ncu3.log 2>&1
tail -2 gpurun_out/r1i_ncu2.log
