#!/bin/bash
# Round 1 (2b): write-combining / row cache of the façade grid, persistent staging of region I/O
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
( time tests/facade/_bin/facade_test ) 2>&1 | tail -12
( time tests/facade/_bin/striping_test ) 2>&1 | tail -6
( time tests/facade/_bin/generic_test ) 2>&1 | tail -6
