#!/bin/bash
# Round 1 (t): slab groups (single-process multi-GPU API) and the C++ striping façade
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_group_gpu.py tests/test_facade_gpu.py -x -q -m gpu > gpurun_out/r1t_pytest.log 2>&1
tail -25 gpurun_out/r1t_pytest.log
tests/facade/_bin/striping_test > gpurun_out/r1t_striping.log 2>&1; echo "striping_test exit $?"; tail -22 gpurun_out/r1t_striping.log
