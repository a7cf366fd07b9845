#!/bin/bash
# GPU job r4g: fused LBM kernel with the ring warp as TMA producer: parity, then tile rows x z chunks x L2 hints at 512^3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lbm_fused_gpu.py -q -m gpu -x > gpurun_out/r4g_pytest.log 2>&1; tail -3 gpurun_out/r4g_pytest.log
timeout 900 python tools/tune.py lbm lbm.tb=2 lbm.tb_rows=14,16,8 lbm.tb_zchunk=32,48,64 lbm.tb_hints=0,3 2>&1 | tee gpurun_out/r4g_tune.log
