#!/bin/bash
mkdir -p gpurun_out
for wl in jacobi7_64 jacobi7_128 jacobi27_128 jacobi7_256; do
timeout 300 python tools/tune.py $wl jacobi.tb=1 jacobi.pdl=0,1 jacobi.zchunk=1,2,4,8 2>&1 | tee -a gpurun_out/r1z2_tune.log
done
timeout 300 python tools/tune.py jacobi7_128 jacobi.tb=2,4 jacobi.pdl=1 jacobi.zchunk=1 2>&1 | tee -a gpurun_out/r1z2_tune.log
