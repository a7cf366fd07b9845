#!/bin/bash
# Round 1 (z): programmatic dependent launch for the one-sweep Jacobi kernel on small grids
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_group_gpu.py -x -q -m gpu -k "jacobi" > gpurun_out/r1z_pytest.log 2>&1; tail -3 gpurun_out/r1z_pytest.log
timeout 300 python tools/tune.py jacobi7_128 jacobi.tb=1 jacobi.pdl=0,1 jacobi.zchunk=2,4,8 > gpurun_out/r1z_tune_j128.log 2>&1; cat gpurun_out/r1z_tune_j128.log
timeout 300 python tools/tune.py jacobi7 jacobi.tb=1 jacobi.pdl=0,1 > gpurun_out/r1z_tune_j7.log 2>&1; cat gpurun_out/r1z_tune_j7.log
