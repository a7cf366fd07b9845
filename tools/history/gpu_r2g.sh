#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 2 -c 1 -o gpurun_out/prof_r2g_tb27 python tools/tune.py jacobi27 jacobi.tb=2 > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log; ls -la gpurun_out/prof_r2g_tb27.ncu-rep
