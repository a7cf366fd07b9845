#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k gol > gpurun_out/r1q_pytest.log 2>&1
tail -5 gpurun_out/r1q_pytest.log
python tools/tune.py gol gol.bits=0,4 gol.bits_rows=0,4,8,16,64 2>&1 | tail -12
