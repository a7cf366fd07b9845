#!/bin/bash
# GPU job r3i: the SM-resident Jacobi kernel for small grids (parity, time per sweep vs the streaming kernel).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "resident or config1 or jacobi_bit_exact" > gpurun_out/r3i_pytest.log 2>&1; tail -4 gpurun_out/r3i_pytest.log
timeout 300 python tools/tune.py jacobi7_128 jacobi.resident=0,1 jacobi.tb=1 2>&1 | tail -3 | tee gpurun_out/r3i_tune.log
timeout 300 python tools/tune.py jacobi7_128 jacobi.resident=0,1 jacobi.tb=1 2>&1 | tail -3 | tee -a gpurun_out/r3i_tune.log
