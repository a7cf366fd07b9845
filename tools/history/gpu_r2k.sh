#!/bin/bash
# Round 1 (2k): the reference's own CUDASimulator (recompiled for sm_100a) beside the hand-written kernels, 512^3 f64
mkdir -p gpurun_out
timeout 600 oracle/_ref/lgd_ref_cuda_jacobi 512 50 2>&1 | tee gpurun_out/r2k_ref_cuda.jsonl
timeout 200 python tools/tune.py jacobi7_512 jacobi.tb=1,4 2>&1 | tail -2 | tee gpurun_out/r2k_ours_512.log
timeout 200 python tools/tune.py jacobi27_512 jacobi.tb=1,2 2>&1 | tail -2 | tee -a gpurun_out/r2k_ours_512.log
