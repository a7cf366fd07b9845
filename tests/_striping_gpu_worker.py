"""Worker for test_multigpu.py: one rank per GPU over NCCL, real engine (libb200geo.so)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from libgeodecomp_b200 import models, synth  # noqa: E402
from libgeodecomp_b200.simulator import SimpleInitializer  # noqa: E402
from libgeodecomp_b200.striping import StripedSimulator, slab_bounds  # noqa: E402
from oracle import oracle_py  # noqa: E402


class SlabInit(SimpleInitializer):
    def __init__(self, members, dims, steps, edge=None):
        SimpleInitializer.__init__(self, dims, steps)
        self.members, self.edge = members, edge

    def grid(self, target):
        o, d = target.boundingBox()
        if self.edge is not None:
            target.setEdge(self.edge)
        sl = tuple(slice(o[i], o[i] + d[i]) for i in reversed(range(len(o))))
        for name, arr in self.members.items():
            target.loadMember(name, arr[sl], origin=o)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    rank, world = dist.get_rank(), dist.get_world_size()
    failures = []

    def check(tag, sim, want, member):
        last = sim.model.dim - 1
        b = slab_bounds(sim.grid.global_dims[last], world)
        mine = sim.getGrid().saveMember(member)
        w = want[b[rank]:b[rank + 1]]
        if mine.shape != w.shape or not np.array_equal(mine.view(np.uint8), np.ascontiguousarray(w).view(np.uint8)):
            failures.append("%s rank %d" % (tag, rank))

    jacobi_cases = [(7, "Cube", 1, 6, (24, 18, 40)), (27, "Cube", 3, 8, (30, 12, 36)),
                    (27, "Torus", 2, 7, (16, 10, 34)), (6, "Torus", 1, 5, (12, 9, 20)),
                    (7, "Torus", 4, 9, (32, 8, 70)), (27, "Cube", 2, 9, (40, 70, 130)), (7, "Torus", 2, 6, (48, 33, 65))]
    # overlap = rim-first schedule with the interior update overlapping the transfer (fused sweeps per
    # round through the temporal-blocked kernel); False = exchange, then step
    for kind, topo, ghost, steps, shape, overlap in [c + (o,) for c in jacobi_cases for o in (True, False)]:
        nz, ny, nx = shape
        data = synth.jacobi_grid(nx, ny, nz, seed=kind)
        model = models.ALL["Jacobi%d%s" % (kind, topo)]
        sim = StripedSimulator(SlabInit({"temp": data}, (nx, ny, nz), steps, edge=0.75), model, rank=rank, world=world,
                               ghost_width=ghost, device=local, dist=dist, overlap=overlap)
        sim.run()
        check("jacobi%d%s g%d overlap %s" % (kind, topo, ghost, overlap), sim, oracle_py.jacobi(kind, topo == "Torus", data, steps, edge=0.75), "temp")

    for topo, ghost, steps, shape in [("Cube", 1, 12, (64, 100)), ("Torus", 3, 10, (48, 70))]:
        ny, nx = shape
        g = synth.gol_grid(nx, ny)
        model = models.ALL["Conway" + topo]
        sim = StripedSimulator(SlabInit({"alive": g}, (nx, ny), steps), model, rank=rank, world=world, ghost_width=ghost,
                               device=local, dist=dist)
        sim.run()
        check("gol%s g%d" % (topo, ghost), sim, oracle_py.gol(topo == "Torus", g, steps), "alive")

    for ghost, steps, shape, overlap in [(1, 9, (20, 12, 16), True), (1, 9, (20, 12, 16), False), (2, 8, (24, 10, 18), True),
                                         (1, 6, (32, 40, 70), True)]:
        nz, ny, nx = shape
        raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
        members = {n: raw[m].view(t) for m, (n, t) in enumerate(models.LBMCellF.members)}
        sim = StripedSimulator(SlabInit(members, (nx, ny, nz), steps), models.LBMCellF, rank=rank, world=world,
                               ghost_width=ghost, device=local, dist=dist, overlap=overlap)
        sim.run()
        want = oracle_py.lbm(raw, steps)
        for m, (n, t) in enumerate(models.LBMCellF.members):
            check("lbm g%d overlap %s %s" % (ghost, overlap, n), sim, want[m].view(t), n)

    # n-body: slabs of containers, one ghost plane of containers (counts + particles) per side
    class CellInit(SimpleInitializer):
        def __init__(self, counts, parts, steps):
            SimpleInitializer.__init__(self, counts.shape[::-1], steps)
            self.counts, self.parts = counts, parts

        def grid(self, target):
            o, d = target.boundingBox()
            sl = tuple(slice(o[i], o[i] + d[i]) for i in reversed(range(3)))
            target.loadCells(self.counts[sl], self.parts[sl], origin=o)

    for real, dims, steps, vel in [(np.float32, (6, 5, 8), 8, 10.0), (np.float64, (4, 4, 12), 6, 12.0)]:
        c, p = synth.nbody_cells(*dims, vel=vel, dtype=real)
        model = (models.NBodyF if real == np.float32 else models.NBodyD).with_params(dt=0.01)
        sim = StripedSimulator(CellInit(c, p, steps), model, rank=rank, world=world, ghost_width=1, device=local, dist=dist)
        sim.run()
        gc, gp = sim.getGrid().saveCells()
        wc, wp = oracle_py.nbody(c, p, steps, dt=0.01)
        b = slab_bounds(dims[2], world)
        if not (np.array_equal(gc, wc[b[rank]:b[rank + 1]]) and
                np.array_equal(gp.view(np.uint8), np.ascontiguousarray(wp[b[rank]:b[rank + 1]]).view(np.uint8))):
            failures.append("nbody %s rank %d" % (np.dtype(real).name, rank))

    with open("%s.%d" % (sys.argv[1], rank), "w") as f:
        f.write("FAIL " + "; ".join(failures) if failures else "OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
