#!/bin/bash
# GPU job r4k: the C++ streamed run (b200streamedrun.h) on the device: parity binary, the other facade binaries, e2e at 1024^3
mkdir -p gpurun_out
timeout 600 tests/facade/_bin/streamed_test 2>&1 | tee gpurun_out/r4k_streamed_test.log | tail -14
timeout 900 python -m pytest tests/test_facade_gpu.py -q -m gpu -x 2>&1 | tail -3
for m in box stream box stream; do timeout 300 tests/facade/_bin/e2e_bench 1024 20 1 $m | cut -c1-330; done 2>&1 | tee gpurun_out/r4k_e2e_cpp.jsonl
for c in 8 32; do echo chunks $c; done
