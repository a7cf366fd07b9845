#!/bin/bash
# GPU job r4p (after r4o): the full-size LBM parity failure of r4_final (6 sweeps, warp-specialized kernel, z chunks of 32): where and how often
timeout 1500 python tools/lbm_race_probe.py 512 6 10 2>&1 | tee gpurun_out/r4p_probe.log | cut -c1-400
