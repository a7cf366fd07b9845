#!/bin/bash
# GPU job r4e: what bounds the fused LBM kernel (ncu --set full, one launch at 512^3)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbm_tb2 -s 1 -c 1 -o gpurun_out/r4e_lbm_tb2 python tools/few_launches.py lbm lbm.tb=2 > /dev/null 2>&1; ls -la gpurun_out/r4e_lbm_tb2.ncu-rep
