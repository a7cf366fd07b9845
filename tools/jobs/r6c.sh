#!/bin/bash
# GPU job r6c: ContainerCell tile layout (values of a tile's neighbourhood staged in shared memory, 16-bit links) beside the
# window layout: parity tests with both, the bench leg with both, ncu --set full of the tile sweep kernel
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_container_gpu.py -m gpu -q > gpurun_out/r6c_pytest.log 2>&1; tail -6 gpurun_out/r6c_pytest.log
timeout 60 python tools/container_bench.py --steps 100 --kernel 1 --no-cpu --no-e2e > gpurun_out/r6c_container_k1.json 2> gpurun_out/r6c_container_k1.err; cut -c 1-700 gpurun_out/r6c_container_k1.json; tail -2 gpurun_out/r6c_container_k1.err
timeout 60 python tools/container_bench.py --steps 100 --kernel 0 --no-cpu --no-e2e > gpurun_out/r6c_container_k0.json 2> gpurun_out/r6c_container_k0.err; cut -c 1-400 gpurun_out/r6c_container_k0.json
timeout 80 ncu --set full --clock-control none --import-source on -k regex:"tile_sweep_kernel|tile_resolve_kernel" -c 3 -o gpurun_out/r6c_container_tiles_full python tools/container_bench.py --steps 20 --kernel 1 --no-cpu --no-e2e --no-verify > /dev/null 2>&1; ls -la gpurun_out/r6c_container_tiles_full.ncu-rep
