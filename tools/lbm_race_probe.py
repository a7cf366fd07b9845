#!/usr/bin/env python
"""Where do the fused LBM kernels differ from one sweep per launch at 512^3? (job r4o) Device-side comparison of all 19
populations after `steps` sweeps, for several tunings, several repetitions each; prints the count and the bounding box /
some coordinates of the differing cells."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Grid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
M = models.LBMCellF
noise = synth.lbm_grid(n, n, 16, noise=0.01, z0=16, nz_total=n)


def run(tuning):
    for k, v in tuning.items():
        capi.set_tuning(k, v)
    grid = B200Grid(M, (n, n, n))
    for z in range(0, n, 16):
        states = synth.lbm_states(n, n, 16, z, n)
        for m, (name, t) in enumerate(M.members):
            grid.loadMember(name, states if name == "state" else noise[m].view(t), origin=(0, 0, z))
    grid.dev.step(M.kernel, steps)
    torch.cuda.synchronize()
    out = []
    for m in range(19):
        t = torch.empty((n, n, n), dtype=torch.float32, device="cuda")
        grid.saveMember(M.members[m][0], out=t, location=capi.CUDA_DEVICE)
        out.append(t.view(torch.int32).clone())
    for k in tuning:
        capi.set_tuning(k, -1)
    del grid
    return out


ref = run({"lbm.tb": 1})
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for tuning in ({"lbm.tb": 2, "lbm.tb_warps": 1, "lbm.tb_zchunk": 64, "lbm.tb_hints": 16}, {"lbm.tb": 2, "lbm.tb_warps": 1, "lbm.tb_zchunk": 64},
               {"lbm.tb": 2, "lbm.tb_warps": 1, "lbm.tb_zchunk": 32}):
    for rep in range(reps):
        got = run(tuning)
        total = 0
        where = None
        for m in range(19):
            d = (got[m] != ref[m])
            c = int(d.sum().item())
            total += c
            if c and where is None:
                idx = d.nonzero()
                where = (M.members[m][0], c, idx.min(0).values.tolist(), idx.max(0).values.tolist(), idx[:6].tolist())
        print(tuning, "rep", rep, "differing values:", total, where, flush=True)
        del got
