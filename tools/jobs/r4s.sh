#!/bin/bash
# GPU job r4s: what the fence and the late arrival in sweep 2 cost (hints 16 = no fence, 32 = sweep-2 arrival before the stores)
timeout 600 python tools/tune.py lbm lbm.tb=2 lbm.tb_warps=1 lbm.tb_zchunk=32 lbm.tb_hints=0,16,32,48 2>&1 | tee gpurun_out/r4s_tune.log
for v in "lbm.tb_hints=16" "lbm.tb_hints=48"; do timeout 900 python tools/lbm_race_probe2.py 100 lbm.tb_zchunk=64 $v 2>&1 | cut -c1-300; done | tee gpurun_out/r4s_probe.log
