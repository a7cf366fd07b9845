#!/bin/bash
# GPU job r4l: what bounds the warp-specialized fused LBM kernel (ncu --set full, one launch at 512^3); whole GPU suite with it as the default
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbm_tb2w -s 1 -c 1 -o gpurun_out/r4l_lbm_tb2w python tools/few_launches.py lbm lbm.tb=2 > /dev/null 2>&1; ls -la gpurun_out/r4l_lbm_tb2w.ncu-rep
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r4l_pytest.log 2>&1; tail -4 gpurun_out/r4l_pytest.log
