#!/bin/bash
# Round 1 (2d): unbound cells on slab groups (b200geo_group_step_with), all façade binaries, group tests
mkdir -p gpurun_out
nvidia-smi -L | wc -l
tests/facade/_bin/generic_test > gpurun_out/r2d_generic.log 2>&1; echo "generic_test exit $?"; tail -24 gpurun_out/r2d_generic.log
tests/facade/_bin/striping_test 2>&1 | tail -2
tests/facade/_bin/facade_test 2>&1 | tail -2
timeout 900 python -m pytest tests/test_group_gpu.py tests/test_facade_gpu.py -x -q -m gpu 2>&1 | tail -2
