#!/usr/bin/env python
"""Times the n-body sweep (re-bin + force kernels) through the C ABI. GPU only.
usage: tools/nbody_bench.py [containers_per_axis=108] [steps=20] [f4|f8]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libgeodecomp_b200 import capi, models, synth


def candidate_pairs(counts):
    """sum over containers of count * (sum of the 27 neighbours' counts): pairs one sweep evaluates"""
    c = counts.astype(np.int64)
    p = np.pad(c, 1)
    hood = np.zeros_like(c)
    nz, ny, nx = c.shape
    for dz in range(3):
        for dy in range(3):
            for dx in range(3):
                hood += p[dz:dz + nz, dy:dy + ny, dx:dx + nx]
    return int((c * hood).sum())


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 108
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    real = np.dtype(sys.argv[3] if len(sys.argv) > 3 else "f4")
    model = (models.NBodyF if real == np.float32 else models.NBodyD)
    if os.environ.get("NBODY_KERNEL"):
        capi.set_tuning("nbody.kernel", int(os.environ["NBODY_KERNEL"]))
    if os.environ.get("NBODY_THREADS"):
        capi.set_tuning("nbody.threads", int(os.environ["NBODY_THREADS"]))
    if os.environ.get("NBODY_RUN"):
        capi.set_tuning("nbody.run", int(os.environ["NBODY_RUN"]))
    t0 = time.time()
    c, p = synth.nbody_cells(n, n, n, dtype=real)
    particles, pairs = int(c.sum()), candidate_pairs(c)
    print("generated %d particles in %d^3 containers (max %d per container) in %.1f s; %.0f candidate pairs per particle"
          % (particles, n, c.max(), time.time() - t0, pairs / particles), flush=True)
    grid = model.grid_class(model, (n, n, n))
    grid.loadCells(c, p)
    params = model.step_params(True)
    grid.dev.step(model.kernel, 3, params=params)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    grid.dev.step(model.kernel, steps, params=params)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    grid.dev.check()
    print("nbody %s n=%d: %.3f ms/step  %.2f G particle updates/s  %.1f G candidate pairs/s" % (
        real.name, n, ms, particles / ms / 1e6, pairs / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
