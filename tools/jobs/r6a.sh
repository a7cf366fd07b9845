#!/bin/bash
# GPU job r6a: the ContainerCell path on the device: parity tests (Python mirror + C++ drop-in), smoke(), the bench leg,
# launch list and ncu --set full of the link-resolution and sweep kernels
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_container_gpu.py tests/test_facade_gpu.py -m gpu -x -q -k "container" > gpurun_out/r6a_pytest.log 2>&1; tail -5 gpurun_out/r6a_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r6a_smoke.log 2>&1; tail -2 gpurun_out/r6a_smoke.log
timeout 200 python tools/container_bench.py --steps 100 > gpurun_out/r6a_container.json 2> gpurun_out/r6a_container.err; tail -c 1500 gpurun_out/r6a_container.json; tail -3 gpurun_out/r6a_container.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r6a_launches.csv python tools/container_bench.py --steps 20 --no-cpu --no-e2e --no-verify > /dev/null 2>&1; wc -l gpurun_out/r6a_launches.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"resolve_kernel|sweep_kernel" -c 3 -o gpurun_out/r6a_container_full python tools/container_bench.py --steps 20 --no-cpu --no-e2e --no-verify > /dev/null 2>&1; ls -la gpurun_out/r6a_container_full.ncu-rep
