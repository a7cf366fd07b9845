#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal" > gpurun_out/r1e_pytest.log 2>&1
tail -3 gpurun_out/r1e_pytest.log
for wl in jacobi27 jacobi7; do
  timeout 600 python tools/tune.py $wl jacobi.tb=2,3 jacobi.tb_rows=31,32,33,34,64 jacobi.tb_zchunk=128,256 >> gpurun_out/r1e_tune.log 2>&1
done
cat gpurun_out/r1e_tune.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 2 -c 1 -o gpurun_out/prof_r1e_tb7_r33 python tools/tune.py jacobi7 jacobi.tb=2 jacobi.tb_rows=33 > gpurun_out/r1e_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_tb -s 2 -c 1 -o gpurun_out/prof_r1e_tb27_r32 python tools/tune.py jacobi27 jacobi.tb=2 jacobi.tb_rows=32 > gpurun_out/r1e_ncu2.log 2>&1
tail -3 gpurun_out/r1e_ncu2.log
