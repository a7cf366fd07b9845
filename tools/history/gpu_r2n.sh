#!/bin/bash
# Round 1 (2n): the 96-byte AoS LBM user cell — reference CUDASimulator vs the generic device path (parity + throughput)
mkdir -p gpurun_out
tests/facade/_bin/generic_test 2>&1 | tail -4
timeout 300 tests/facade/_bin/generic_test --bench 2>&1 | tee gpurun_out/r2n_generic_bench.jsonl
timeout 300 oracle/_ref/lgd_ref_cuda_jacobi 256 20 2>&1 | tee gpurun_out/r2n_ref_cuda_256.jsonl
