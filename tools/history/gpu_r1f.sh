#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal" > gpurun_out/r1f_pytest.log 2>&1
tail -3 gpurun_out/r1f_pytest.log
for wl in jacobi27 jacobi7; do
  timeout 600 python tools/tune.py $wl jacobi.tb=2 jacobi.tb_rows=32,33,34,64 >> gpurun_out/r1f_tune.log 2>&1
done
cat gpurun_out/r1f_tune.log
