#!/bin/bash
# Round 1 (2m): compute-sanitizer racecheck of the shared-memory kernels (temporal-blocked Jacobi, fused n-body),
# memcheck of the n-body and slab-group paths; small cases only
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "z_chunks" > gpurun_out/r2m_racecheck_tb.log 2>&1; echo "racecheck tb exit $?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r2m_racecheck_tb.log | tail -3
timeout 100 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_nbody_gpu.py -x -q -m gpu -k "config5 or other_capacity" > gpurun_out/r2m_racecheck_nbody.log 2>&1; echo "racecheck nbody exit $?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r2m_racecheck_nbody.log | tail -3
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_nbody_gpu.py tests/test_group_gpu.py -x -q -m gpu -k "dense or boxgroup or group_lbm or group_gol" > gpurun_out/r2m_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r2m_memcheck.log | tail -3
