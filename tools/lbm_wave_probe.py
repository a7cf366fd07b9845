#!/usr/bin/env python
"""One fused LBM launch on a grid of a given shape — the target of `ncu --metrics dram__bytes_*` when asking how much of the
fused kernel's read over-fetch is CTAs drifting apart (a grid of <= 148 tiles starts all its CTAs at the same time).
usage: tools/lbm_wave_probe.py nx ny nz [key=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Grid

nx, ny, nz = (int(a) for a in sys.argv[1:4])
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    capi.set_tuning(k, int(v))
grid = B200Grid(models.LBMCellF, (nx, ny, nz))
grid.loadMember("C", np.ones((nz, ny, nx), dtype=np.float32))
grid.loadMember("density", np.ones((nz, ny, nx), dtype=np.float32))
grid.loadMember("state", synth.lbm_states(nx, ny, nz, 0, nz))
torch.cuda.synchronize()
grid.dev.step(capi.KERNEL_LBM_D3Q19, 8)
torch.cuda.synchronize()
print("compulsory bytes per launch: read %.1f MB, write %.1f MB" % (20 * 4 * nx * ny * nz / 1e6, 19 * 4 * nx * ny * nz / 1e6))
