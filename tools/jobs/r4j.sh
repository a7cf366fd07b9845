#!/bin/bash
# GPU job r4j: fused LBM kernel, two CTAs per SM with more useful rows (one TMA stage) vs one CTA per SM
timeout 600 python tools/tune.py lbm lbm.tb=2 lbm.tb_rows=14,8,92,111,121 lbm.tb_zchunk=32,64 2>&1 | tee gpurun_out/r4j_tune.log
