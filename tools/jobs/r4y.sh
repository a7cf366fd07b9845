#!/bin/bash
# GPU job r4y: warp-specialized LBM kernel with L2 prefetch of the TMA windows (lbm.tb_prefetch = planes ahead)
timeout 600 python tools/tune.py lbm lbm.tb=2 lbm.tb_prefetch=0,1,2,4,8 lbm.tb_rows=123,142 2>&1 | tee gpurun_out/r4y_tune.log
timeout 300 python tools/lbm_race_probe2.py 60 lbm.tb_prefetch=4 2>&1 | cut -c1-300 | tee gpurun_out/r4y_probe.log
