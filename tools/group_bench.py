#!/usr/bin/env python
"""Weak-scaling throughput of the single-process multi-GPU path (b200geo_group_*: one host thread, one
slab per GPU, rim-first schedule, direct NVLink copies between ghost planes). GPU only.
usage: tools/group_bench.py jacobi27|jacobi7|lbm [n_gpus] [ghost_width] [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libgeodecomp_b200 import capi, models
from libgeodecomp_b200.simulator import B200Grid

SPEC = {"jacobi27": (models.Jacobi27Cube, (1024, 1024, 1024), 2), "jacobi7": (models.Jacobi7Cube, (1024, 1024, 1024), 2),
        "lbm": (models.LBMCellF, (512, 512, 512), 1)}


def main():
    wl = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
    model, dims, w = SPEC[wl]
    w = int(sys.argv[3]) if len(sys.argv) > 3 else w
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    capi.set_tuning("jacobi.tb", w)
    gdims = (dims[0], dims[1], dims[2] * n)
    grids = []
    plane = np.random.default_rng(0).random((1, dims[1], dims[0]))
    for r in range(n):
        modes = [capi.GHOST_PEER if r > 0 else capi.GHOST_EDGE, capi.GHOST_PEER if r < n - 1 else capi.GHOST_EDGE]
        if n == 1:
            g = B200Grid(model, dims, device=r)
        else:
            g = B200Grid(model, dims, device=r, ghost_z=w, z_modes=modes, origin=[0, 0, r * dims[2]], global_dims=gdims)
        if wl == "lbm":
            from libgeodecomp_b200 import synth
            g.loadMember("C", np.ones(dims[::-1], dtype=np.float32))
            g.loadMember("density", np.ones(dims[::-1], dtype=np.float32))
            g.loadMember("state", synth.lbm_states(dims[0], dims[1], dims[2], r * dims[2], gdims[2]))
        else:
            block = np.broadcast_to(plane, (32, dims[1], dims[0]))
            for z in range(0, dims[2], 32):
                g.loadMember("temp", block, origin=[0, 0, r * dims[2] + z])
        grids.append(g)
    group = capi.SlabGroup([g.dev for g in grids])
    group.step(model.kernel, 2 * w + 4)
    group.sync()
    for d in range(n):
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    group.step(model.kernel, steps)
    group.sync()
    dt = time.perf_counter() - t0
    cells = float(np.prod(dims)) * n
    st = group.stats()
    print(json.dumps({"workload": wl, "path": "b200geo_group (one process, %d GPU(s))" % n, "n_gpus": n, "ghost_width": w,
                      "steps": steps, "ms_per_step": 1e3 * dt / steps, "glups": 1e-9 * cells * steps / dt,
                      "exchanges": st["exchanges"], "halo_bytes": st["bytes"], "timing": "host wall clock around step + sync"}))


if __name__ == "__main__":
    main()
