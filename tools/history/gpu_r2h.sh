#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal_blocking or z_chunks" 2>&1 | tail -2
timeout 300 python tools/tune.py jacobi27 jacobi.tb=2 jacobi.tb_rows=33 2>&1 | tail -1
timeout 300 python tools/tune.py jacobi27 jacobi.tb=2 jacobi.tb_rows=33 2>&1 | tail -1
timeout 300 python tools/tune.py jacobi7 jacobi.tb=4 jacobi.tb_rows=33 2>&1 | tail -1
