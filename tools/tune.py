#!/usr/bin/env python
"""Times kernel variants through the C ABI (b200geo_set_tuning). GPU only.
usage: tools/tune.py jacobi27|jacobi7|lbm|gol key=v1,v2,... [key=...]"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Grid

SPEC = {"jacobi27": (models.Jacobi27Cube, (1024, 1024, 1024), 16), "jacobi7": (models.Jacobi7Cube, (1024, 1024, 1024), 16),
        "lbm": (models.LBMCellF, (512, 512, 512), 152), "gol": (models.ConwayCube, (16384, 16384), 2),
        "jacobi7_128": (models.Jacobi7Cube, (128, 128, 128), 16), "jacobi7_64": (models.Jacobi7Cube, (64, 64, 64), 16),
        "jacobi7_256": (models.Jacobi7Cube, (256, 256, 256), 16), "jacobi27_128": (models.Jacobi27Cube, (128, 128, 128), 16),
        "jacobi7_512": (models.Jacobi7Cube, (512, 512, 512), 16), "jacobi27_512": (models.Jacobi27Cube, (512, 512, 512), 16)}


def main():
    wl = sys.argv[1]
    model, dims, nbytes = SPEC[wl]
    sweeps = {k: [int(x) for x in v.split(",")] for k, v in (a.split("=") for a in sys.argv[2:])}
    grid = B200Grid(model, dims)
    if wl == "gol":
        a = (np.random.default_rng(0).random((dims[1], dims[0])) < 0.35).astype(np.uint8)
        grid.loadMember("alive", a)
    elif wl == "lbm":
        # lid-driven cavity: fluid at rest inside the six walls of the reference example
        grid.loadMember("C", np.ones(dims[::-1], dtype=np.float32))
        grid.loadMember("density", np.ones(dims[::-1], dtype=np.float32))
        grid.loadMember("state", synth.lbm_states(dims[0], dims[1], dims[2], 0, dims[2]))
    else:
        grid.loadMember("temp", np.random.default_rng(0).random(dims[::-1]))
    cells = float(np.prod(dims))
    keys = sorted(sweeps)
    for combo in itertools.product(*[sweeps[k] for k in keys]):
        for k, v in zip(keys, combo):
            capi.set_tuning(k, v)
        grid.dev.step(model.kernel, 5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        grid.dev.step(model.kernel, n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("%s %s: %.4f ms/sweep  %.1f GLUPS  %.0f GB/s algorithmic" % (
            wl, dict(zip(keys, combo)), ms, cells / ms / 1e6, nbytes * cells / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
