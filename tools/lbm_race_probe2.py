#!/usr/bin/env python
"""ONE fused LBM launch (two sweeps) at 512^3, many times from the same input, each result compared on the device with the
result of two single-sweep launches: how often, where and in which populations the warp-specialized kernel goes wrong
(job r4q). usage: tools/lbm_race_probe2.py reps key=value ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Grid

n = 512
reps = int(sys.argv[1])
tuning = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in sys.argv[2:])
M = models.LBMCellF
noise = synth.lbm_grid(n, n, 16, noise=0.01, z0=16, nz_total=n)


def fresh():
    grid = B200Grid(M, (n, n, n))
    for z in range(0, n, 16):
        states = synth.lbm_states(n, n, 16, z, n)
        for m, (name, t) in enumerate(M.members):
            grid.loadMember(name, states if name == "state" else noise[m].view(t), origin=(0, 0, z))
    return grid


def pops(grid):
    out = []
    for m in range(19):
        t = torch.empty((n, n, n), dtype=torch.float32, device="cuda")
        grid.saveMember(M.members[m][0], out=t, location=capi.CUDA_DEVICE)
        out.append(t.view(torch.int32))
    return out


capi.set_tuning("lbm.tb", 1)
g = fresh()
g.dev.step(M.kernel, 2)
ref = pops(g)
del g
capi.set_tuning("lbm.tb", 2)
for k, v in tuning.items():
    capi.set_tuning(k, v)
g = fresh()
bad_runs = 0
for rep in range(reps):
    g.dev.step(M.kernel, 2)      # buffers: cur -> other; the input buffer is untouched (both hold what a launch needs)
    got = pops(g)
    total, report = 0, []
    for m in range(19):
        d = got[m] != ref[m]
        c = int(d.sum().item())
        total += c
        if c:
            idx = d.nonzero()
            report.append((M.members[m][0], c, idx.min(0).values.tolist(), idx.max(0).values.tolist()))
    if total:
        bad_runs += 1
        print("rep", rep, "differing values", total, report[:20], flush=True)
    del got
    g.dev.swap()                 # back to the input buffer: the next launch starts from the same state
print(tuning, "launches with errors:", bad_runs, "of", reps, flush=True)
