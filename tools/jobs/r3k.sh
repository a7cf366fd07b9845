#!/bin/bash
# GPU job r3k: what bounds the SM-resident kernel (ncu --set full, one launch of 50 sweeps); C++ e2e again.
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resident -c 1 -o gpurun_out/r3k_resident python tools/few_launches.py jacobi7_128 jacobi.tb=1 --sweeps 50 > /dev/null 2>&1; ls -la gpurun_out/r3k_resident.ncu-rep
for mode in box rows; do timeout 300 tests/facade/_bin/e2e_bench 1024 20 1 $mode | cut -c1-260; done | tee gpurun_out/r3k_e2e_cpp.jsonl
