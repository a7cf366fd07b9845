#!/bin/bash
# GPU job r4v: compute-sanitizer on the code of this session: fused LBM kernels (memcheck, racecheck), the C++ streamed run and
# B200PatchLink (memcheck)
mkdir -p gpurun_out
K='test_lbm_fused_bit_exact and (shape1 or shape2 or shape4) and not 0-16 and not 0-8 or walls_anywhere and 1-123 or update_box_two_sweeps and origin1'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_lbm_fused_gpu.py -q -m gpu -x -k "$K" > gpurun_out/r4v_memcheck_lbm.log 2>&1; echo "memcheck lbm rc=$?"; tail -4 gpurun_out/r4v_memcheck_lbm.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_lbm_fused_gpu.py -q -m gpu -x -k "test_lbm_fused_bit_exact and shape4 and (1-123 or 0-14) and 8-" > gpurun_out/r4v_racecheck_lbm.log 2>&1; echo "racecheck lbm rc=$?"; tail -6 gpurun_out/r4v_racecheck_lbm.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 tests/facade/_bin/streamed_test > gpurun_out/r4v_memcheck_streamed.log 2>&1; echo "memcheck streamed rc=$?"; tail -3 gpurun_out/r4v_memcheck_streamed.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 tests/facade/_bin/stepper_test > gpurun_out/r4v_memcheck_stepper.log 2>&1; echo "memcheck stepper rc=$?"; tail -3 gpurun_out/r4v_memcheck_stepper.log
