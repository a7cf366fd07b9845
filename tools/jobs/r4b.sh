#!/bin/bash
# GPU job r4b: fused LBM kernel, second version (sweep-1 input through TMA windows): parity, tile shapes / stages / z chunks at 512^3.
mkdir -p gpurun_out
for i in 0 1 2 3 4 5; do tools/probe/tma_probe $i; done 2>&1 | tee gpurun_out/r4b_tma_probe.log
timeout 900 python -m pytest tests/test_lbm_fused_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x -k "lbm" > gpurun_out/r4b_pytest.log 2>&1; tail -8 gpurun_out/r4b_pytest.log
timeout 600 python tools/tune.py lbm lbm.tb=1 > gpurun_out/r4b_tune.log 2>&1
timeout 900 python tools/tune.py lbm lbm.tb=2 lbm.tb_rows=14,142,16,8 lbm.tb_zchunk=32,64,128 >> gpurun_out/r4b_tune.log 2>&1
timeout 900 python tools/tune.py lbm lbm.tb=2 lbm.tb_rows=14 lbm.tb_promo=1,2,3 >> gpurun_out/r4b_tune.log 2>&1
cat gpurun_out/r4b_tune.log
