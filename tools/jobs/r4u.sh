#!/bin/bash
# GPU job r4u: warp-specialized LBM kernel, rows x stages x z chunks; the default (12 x 3) under the single-launch probe; all LBM tests
timeout 600 python tools/tune.py lbm lbm.tb=2 lbm.tb_warps=1 lbm.tb_rows=123,113,104,103 lbm.tb_zchunk=16,32,64 2>&1 | tee gpurun_out/r4u_tune.log
for v in "lbm.tb_zchunk=32" "lbm.tb_zchunk=64" "lbm.tb_zchunk=32 lbm.tb_rows=104"; do timeout 900 python tools/lbm_race_probe2.py 150 $v 2>&1 | cut -c1-300; done | tee gpurun_out/r4u_probe.log
timeout 600 python -m pytest tests/test_lbm_fused_gpu.py tests/test_fullsize_gpu.py tests/test_streamed_gpu.py tests/test_group_gpu.py -q -m gpu -x -k "lbm" 2>&1 | tail -3
