#!/bin/bash
# GPU job r6d: the tile layout after the staging table moved to one thread per staged container: bench leg, verified
mkdir -p gpurun_out
timeout 35 python tools/container_bench.py --steps 100 --kernel 1 --no-cpu --no-e2e > gpurun_out/r6d_container_k1.json 2> gpurun_out/r6d_container_k1.err; cut -c 1-900 gpurun_out/r6d_container_k1.json; tail -2 gpurun_out/r6d_container_k1.err
