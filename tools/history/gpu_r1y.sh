#!/bin/bash
# Round 1 (y): n-body with the candidate-parallel filter (nbody.kernel = 4): parity of every variant, timing 3 vs 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nbody_gpu.py -x -q -m gpu > gpurun_out/r1y_pytest.log 2>&1; tail -5 gpurun_out/r1y_pytest.log
for k in 3 4; do
  NBODY_KERNEL=$k timeout 300 python tools/nbody_bench.py 108 10 f4 2>&1 | tail -2
done
NBODY_KERNEL=3 timeout 300 python tools/nbody_bench.py 64 6 f8 2>&1 | tail -1
NBODY_KERNEL=4 timeout 300 python tools/nbody_bench.py 64 6 f8 2>&1 | tail -1
