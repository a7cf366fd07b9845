#!/bin/bash
# GPU job r4i: fused LBM kernel, where the time goes: TMA windows only (hints 4: no compute, no stores), sweep 1 only (hints 8: no sweep 2)
timeout 600 python tools/tune.py lbm lbm.tb=2 lbm.tb_hints=0,4,8 lbm.tb_rows=14,8 2>&1 | tee gpurun_out/r4i_tune.log
