#!/bin/bash
# GPU job r5g: ncu --set full of the 7-point T = 4 Jacobi kernel and of the n-body kernel of the final build
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 1 -c 1 -o gpurun_out/r5g_tb7_full python tools/few_launches.py jacobi7 jacobi.tb=4 --sweeps 16 > /dev/null 2>&1; ls -la gpurun_out/r5g_tb7_full.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/r5g_nbody_full python tools/nbody_bench.py 108 4 f4 > /dev/null 2>&1; ls -la gpurun_out/r5g_nbody_full.ncu-rep
