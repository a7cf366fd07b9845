#!/bin/bash
# round 1d: temporal-blocked Jacobi — parity, then a sweep over depth / tile shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal or jacobi" > gpurun_out/r1d_pytest.log 2>&1
tail -5 gpurun_out/r1d_pytest.log
for wl in jacobi27 jacobi7; do
  timeout 600 python tools/tune.py $wl jacobi.tb=1,2,3,4 jacobi.tb_rows=32,33,64 >> gpurun_out/r1d_tune.log 2>&1
done
timeout 300 python tools/tune.py jacobi27 jacobi.tb=2,4 jacobi.tb_rows=32 jacobi.tb_zchunk=32,64,256,1024 >> gpurun_out/r1d_tune.log 2>&1
cat gpurun_out/r1d_tune.log
