#!/bin/bash
# Round 1 (2f): slab groups of BoxCell containers (n-body on several GPUs from one process), C++ façade
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_group_gpu.py -x -q -m gpu -k "boxgroup or rejects" 2>&1 | tail -4
tests/facade/_bin/striping_test > gpurun_out/r2f_striping.log 2>&1; echo "striping_test exit $?"; tail -12 gpurun_out/r2f_striping.log
tests/facade/_bin/facade_test 2>&1 | tail -2
tests/facade/_bin/generic_test 2>&1 | tail -1
