#!/bin/bash
# GPU job r4k: the C++ streamed run (b200streamedrun.h) on the device: parity binary, the other facade binaries, e2e at 1024^3;
# the warp-specialized fused LBM kernel (lbm.tb_warps = 1): parity under a short timeout (a lost mbarrier arrival would hang), tuning
mkdir -p gpurun_out
timeout 600 tests/facade/_bin/streamed_test 2>&1 | tee gpurun_out/r4k_streamed_test.log | tail -14
timeout 900 python -m pytest tests/test_facade_gpu.py -q -m gpu -x 2>&1 | tail -3
for m in box stream box stream; do timeout 300 tests/facade/_bin/e2e_bench 1024 20 1 $m | cut -c1-330; done 2>&1 | tee gpurun_out/r4k_e2e_cpp.jsonl
timeout 120 python -m pytest tests/test_lbm_fused_gpu.py -q -m gpu -x -k "1-14 or 1-12 or 1-123 or 1-10 or warps1 or 1]" > gpurun_out/r4k_pytest_warps.log 2>&1; echo "pytest warps rc=$?"; tail -5 gpurun_out/r4k_pytest_warps.log
timeout 300 python -m pytest tests/test_lbm_fused_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x -k "lbm" > gpurun_out/r4k_pytest_lbm.log 2>&1; echo "pytest lbm rc=$?"; tail -3 gpurun_out/r4k_pytest_lbm.log
timeout 300 python tools/tune.py lbm lbm.tb=2 lbm.tb_warps=0,1 2>&1 | tee gpurun_out/r4k_tune.log
timeout 300 python tools/tune.py lbm lbm.tb=2 lbm.tb_warps=1 lbm.tb_rows=14,12,123,10 lbm.tb_zchunk=32,64,128 2>&1 | tee -a gpurun_out/r4k_tune.log
