#!/bin/bash
# GPU job r4a: the LBM kernel that fuses two sweeps (csrc/lbm_tb.cu): parity, tile shapes / z chunks at 512^3; the resident Jacobi kernel with precomputed shell offsets.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -2
timeout 900 python -m pytest tests/test_lbm_fused_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x -k "lbm or resident" > gpurun_out/r4a_pytest.log 2>&1; tail -8 gpurun_out/r4a_pytest.log
timeout 600 python tools/tune.py lbm lbm.tb=1 > gpurun_out/r4a_tune.log 2>&1
timeout 900 python tools/tune.py lbm lbm.tb=2 lbm.tb_rows=14,16 lbm.tb_zchunk=32,64,128,256 >> gpurun_out/r4a_tune.log 2>&1
timeout 300 python tools/tune.py jacobi7_128 jacobi.resident=0,1 >> gpurun_out/r4a_tune.log 2>&1
cat gpurun_out/r4a_tune.log
