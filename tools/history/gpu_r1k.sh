#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "2" > gpurun_out/r1k_pytest.log 2>&1
tail -5 gpurun_out/r1k_pytest.log
timeout 900 python bench.py --steps 100 --warmup 6 > gpurun_out/r1k_bench_n1.json 2> gpurun_out/r1k_bench_n1.err
tail -c 2500 gpurun_out/r1k_bench_n1.json; tail -3 gpurun_out/r1k_bench_n1.err
