#!/bin/bash
# Round 1 (s): full GPU test suite (incl. the generic device path and the LBM variants), then tuning
# sweeps of the LBM variants and of the temporal-blocking depth / tile shape of the Jacobi kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r1s_pytest.log 2>&1
tail -5 gpurun_out/r1s_pytest.log
tests/facade/_bin/generic_test > gpurun_out/r1s_generic.log 2>&1; echo "generic_test exit $?"; tail -14 gpurun_out/r1s_generic.log
timeout 300 python tools/tune.py lbm lbm.variant=1,2 lbm.block=128,256 > gpurun_out/r1s_tune_lbm.log 2>&1; cat gpurun_out/r1s_tune_lbm.log | tail -8
timeout 400 python tools/tune.py jacobi27 jacobi.tb=2,3,4 jacobi.tb_rows=31,32,33,34,64 > gpurun_out/r1s_tune_j27.log 2>&1; tail -16 gpurun_out/r1s_tune_j27.log
timeout 300 python tools/tune.py jacobi7 jacobi.tb=2,3,4 jacobi.tb_rows=32,33,34 > gpurun_out/r1s_tune_j7.log 2>&1; tail -10 gpurun_out/r1s_tune_j7.log
timeout 200 python tools/tune.py jacobi7_128 jacobi.tb=1 jacobi.zchunk=2,4,8,16,32 jacobi.prefetch=0,2 > gpurun_out/r1s_tune_j128.log 2>&1; tail -10 gpurun_out/r1s_tune_j128.log
