#!/bin/bash
# GPU job r4w (as r4t): the tma_empty arrival behind an instruction that needs every loaded value: 300 single launches, LBM tests, speed
# a word first): 300 single launches, parity tests, speed
for v in "lbm.tb_zchunk=64" "lbm.tb_zchunk=32"; do timeout 900 python tools/lbm_race_probe2.py 150 $v 2>&1 | cut -c1-300; done | tee gpurun_out/r4w_probe.log
timeout 600 python -m pytest tests/test_lbm_fused_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x -k lbm 2>&1 | tail -3
timeout 300 python tools/tune.py lbm lbm.tb=2 lbm.tb_warps=0,1 lbm.tb_zchunk=32,64 2>&1 | tee gpurun_out/r4w_tune.log
timeout 300 python tools/tune.py lbm lbm.tb=2 lbm.tb_warps=1 lbm.tb_rows=142 lbm.tb_zchunk=32,64 2>&1 | tee -a gpurun_out/r4w_tune.log
