"""Generates tests/golden/*.npz by running the REFERENCE's own SerialSimulator (oracle/_ref/lgd_ref_*,
built from /root/reference by oracle/Makefile) on small seeded inputs. Run in the build container:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures hold input and output so the tests do not depend on the generator staying unchanged.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libgeodecomp_b200 import synth  # noqa: E402
from oracle import oracle_py  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    cases = {}
    for kind in (6, 7, 27):
        for topo in ("cube", "torus"):
            for (nx, ny, nz, steps) in [(12, 10, 8, 5), (33, 5, 7, 3)]:
                g = synth.jacobi_grid(nx, ny, nz, seed=100 + kind)
                edge = 0.5 if topo == "cube" else None
                out, _ = oracle_py.run_ref("jacobi%d%s" % (kind, topo), g, (nx, ny, nz), steps, edge=edge)
                key = "jacobi%d_%s_%dx%dx%d_s%d" % (kind, topo, nx, ny, nz, steps)
                cases[key + "_in"] = g
                cases[key + "_out"] = out.view(np.float64).reshape(g.shape)
    np.savez_compressed(os.path.join(HERE, "jacobi.npz"), **cases)

    cases = {}
    for topo in ("cube", "torus"):
        for (nx, ny, steps) in [(40, 30, 12), (17, 9, 30), (160, 90, 40)]:
            g = synth.gol_grid(nx, ny, seed=3)
            out, _ = oracle_py.run_ref("conway" + topo, g, (nx, ny, 1), steps)
            key = "gol_%s_%dx%d_s%d" % (topo, nx, ny, steps)
            cases[key + "_in"] = g
            cases[key + "_out"] = out.reshape(g.shape)
    np.savez_compressed(os.path.join(HERE, "gol.npz"), **cases)

    cases = {}
    for (nx, ny, nz, steps) in [(10, 8, 6, 9), (16, 5, 4, 4)]:
        raw = synth.lbm_grid(nx, ny, nz, noise=0.02)
        out, _ = oracle_py.run_ref("lbm", raw, (nx, ny, nz), steps)
        key = "lbm_%dx%dx%d_s%d" % (nx, ny, nz, steps)
        cases[key + "_in"] = raw
        cases[key + "_out"] = out.view(np.float32).reshape(raw.shape)
    np.savez_compressed(os.path.join(HERE, "lbm.npz"), **cases)
    cases = {}
    for real, (ncx, ncy, ncz, steps, vel, dt) in [(np.float32, (6, 5, 4, 10, 8.0, 0.01)), (np.float64, (6, 5, 4, 10, 8.0, 0.01)),
                                                  (np.float32, (9, 3, 2, 6, 20.0, 0.02)), (np.float64, (3, 3, 3, 25, 2.0, 0.005))]:
        c, p = synth.nbody_cells(ncx, ncy, ncz, vel=vel, dtype=real)
        (co, po), _ = oracle_py.run_ref_nbody(c, p, steps, dt=dt)
        key = "nbody_%s_%dx%dx%d_s%d_dt%g" % (np.dtype(real).name, ncx, ncy, ncz, steps, dt)
        cases[key + "_in_counts"], cases[key + "_in_parts"] = c, p
        cases[key + "_out_counts"], cases[key + "_out_parts"] = co, po
    np.savez_compressed(os.path.join(HERE, "nbody.npz"), **cases)
    # ContainerCell grids of ID-keyed mesh elements (oracle/models/container.h): capacity 16, 20 neighbour ids.
    # The last two cases hold the SAME id in several containers (first hit in the adapter's search order wins).
    cases = {}
    for key, (dims, nd, torus, edge, steps, dup) in {
            "container_3cube_6x5x4_s9": ((6, 5, 4), 3, False, False, 9, False), "container_3torus_6x5x4_s9": ((6, 5, 4), 3, True, False, 9, False),
            "container_3cube_edge_5x4x3_s6": ((5, 4, 3), 3, False, True, 6, False), "container_2cube_edge_9x7_s12": ((9, 7, 1), 2, False, True, 12, False),
            "container_2torus_9x7_s12": ((9, 7, 1), 2, True, False, 12, False), "container_3torus_2x1x3_s5": ((2, 1, 3), 3, True, False, 5, False),
            "container_3cube_dup_4x4x3_s4": ((4, 4, 3), 3, False, True, 4, True), "container_2torus_dup_5x5_s4": ((5, 5, 1), 2, True, False, 4, True)}.items():
        box, eb = synth.container_cells(*dims, n_dims=nd, torus=torus, edge=edge, seed=500 + len(cases))
        if dup:
            synth.container_duplicate_ids(box)
            if eb is not None:
                synth.container_duplicate_ids(eb)
        out, _ = oracle_py.run_ref_container(box, steps, n_dims=nd, torus=torus, edge=eb)
        for n in oracle_py.CONTAINER_FIELDS:
            cases[key + "_in_" + n] = box[n]
            if eb is not None:
                cases[key + "_edge_" + n] = eb[n]
        cases[key + "_out_values"] = out["values"]
    np.savez_compressed(os.path.join(HERE, "container.npz"), **cases)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
