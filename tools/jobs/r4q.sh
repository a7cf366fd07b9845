#!/bin/bash
# GPU job r4q: which dependency of the warp-specialized LBM kernel is violated: single launches from the same input, 60 each
for v in "lbm.tb_zchunk=64" "lbm.tb_zchunk=64 lbm.tb_hints=32" "lbm.tb_zchunk=64 lbm.tb_hints=64" "lbm.tb_zchunk=64 lbm.tb_warps=0"; do
  timeout 600 python tools/lbm_race_probe2.py 60 $v 2>&1 | cut -c1-700
done | tee gpurun_out/r4q_probe.log
