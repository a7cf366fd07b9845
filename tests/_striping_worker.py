"""Worker for test_striping_gloo.py: one rank of a world_size-N gloo job on CPU."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpu_engine  # noqa: E402
from libgeodecomp_b200 import models, synth  # noqa: E402
from libgeodecomp_b200.simulator import SimpleInitializer  # noqa: E402
from libgeodecomp_b200.striping import StripedSimulator, slab_bounds  # noqa: E402
from oracle import oracle_py  # noqa: E402


class SlabInit(SimpleInitializer):
    """initialises only the cells inside target.boundingBox() (io/initializer.h:38-44)"""

    def __init__(self, data, steps, edge):
        SimpleInitializer.__init__(self, data.shape[::-1], steps)
        self.data, self.edge = data, edge

    def grid(self, target):
        (ox, oy, oz), (dx, dy, dz) = target.boundingBox()
        target.setEdge(self.edge)
        target.loadMember("temp", self.data[oz:oz + dz, oy:oy + dy, ox:ox + dx], origin=(ox, oy, oz))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    out_path = sys.argv[1]
    failures = []
    cases = [(7, 1, 5, (12, 9, 10)), (27, 3, 7, (16, 6, 8)), (6, 2, 6, (13, 5, 4)), (27, 4, 9, (12, 4, 4))]
    # every case with the rim-first / interior-overlapped schedule and with exchange-then-step
    for kind, ghost, steps, shape, overlap in [c + (o,) for c in cases for o in (True, False)] + [(7, 2, 8, (24, 5, 6), True)]:
        nz, ny, nx = shape
        data = synth.jacobi_grid(nx, ny, nz, seed=kind)
        model = models.ALL["Jacobi%dCube" % kind]
        sim = StripedSimulator(SlabInit(data, steps, 0.75), model, rank=rank, world=world, ghost_width=ghost,
                               dist=dist, engine=cpu_engine, overlap=overlap)
        sim.run()
        assert sim.getStep() == steps
        b = slab_bounds(nz, world)
        mine = sim.getGrid().saveMember("temp")
        want = oracle_py.jacobi(kind, False, data, steps, edge=0.75)[b[rank]:b[rank + 1]]
        if mine.shape != want.shape or not np.array_equal(mine, want):
            failures.append("kind %d ghost %d overlap %s rank %d" % (kind, ghost, overlap, rank))
    # Torus: the slab ring is closed (rank 0's low neighbour is the last rank; with two ranks both neighbours are the
    # same rank and the messages are told apart by posting order)
    for kind, ghost, steps, shape, overlap in [(27, 2, 7, (12, 5, 6), True), (6, 1, 5, (9, 4, 7), True), (7, 3, 8, (18, 4, 5), False),
                                               (27, 1, 4, (6, 6, 6), False)]:
        nz, ny, nx = shape
        data = synth.jacobi_grid(nx, ny, nz, seed=kind + 100)
        model = models.ALL["Jacobi%dTorus" % kind]
        sim = StripedSimulator(SlabInit(data, steps, 0.0), model, rank=rank, world=world, ghost_width=ghost,
                               dist=dist, engine=cpu_engine, overlap=overlap)
        sim.run()
        b = slab_bounds(nz, world)
        mine = sim.getGrid().saveMember("temp")
        want = oracle_py.jacobi(kind, True, data, steps)[b[rank]:b[rank + 1]]
        if mine.shape != want.shape or not np.array_equal(mine, want):
            failures.append("torus kind %d ghost %d overlap %s rank %d" % (kind, ghost, overlap, rank))

    # LBM D3Q19 (24 members): ghost width 1 ships only the populations that cross the face, ghost width 2 (two fused sweeps
    # per round, the rim recomputed) the 19 populations and the state — member by member and in place, with the rim-first
    # schedule; the last case forces the packed transport (whole cells through saveRegion / loadRegion, exchange first,
    # then step: what a multi-member model without a halo_members() list gets)
    class LBMInit(SimpleInitializer):
        def __init__(self, raw, steps):
            SimpleInitializer.__init__(self, raw.shape[1:][::-1], steps)
            self.raw = raw

        def grid(self, target):
            (ox, oy, oz), (dx, dy, dz) = target.boundingBox()
            for m, (name, t) in enumerate(models.LBMCellF.members):
                target.loadMember(name, self.raw[m, oz:oz + dz, oy:oy + dy, ox:ox + dx].view(t), origin=(ox, oy, oz))

    for ghost, steps, overlap, packed in ((1, 5, True, False), (1, 4, False, False), (2, 5, True, False), (2, 4, False, False),
                                         (2, 5, True, True)):
        nz, ny, nx = 12, 6, 7
        raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
        sim = StripedSimulator(LBMInit(raw, steps), models.LBMCellF, rank=rank, world=world, ghost_width=ghost,
                               dist=dist, engine=cpu_engine, overlap=overlap)
        if packed:
            sim.halo.packed = True
            sim.halo.need = (list(range(24)), list(range(24)))
        elif sim.halo.packed or (overlap and not sim._can_overlap()):
            failures.append("lbm ghost %d: halos are packed / rounds not overlapped" % ghost)
        sim.run()
        b = slab_bounds(nz, world)
        want = oracle_py.lbm(raw, steps)
        for m, (name, t) in enumerate(models.LBMCellF.members):
            mine = sim.getGrid().saveMember(name)
            if not np.array_equal(mine.view(np.int32), want[m, b[rank]:b[rank + 1]].view(np.int32)):
                failures.append("lbm ghost %d overlap %s member %s rank %d" % (ghost, overlap, name, rank))
                break
    # ParallelWriters are handed the rank's slab as validRegion, the global dimensions and the rank (io/parallelwriter.h:92-99);
    # the same writer pulls its slab — this is what bench.py's e2e leg does on every rank
    from libgeodecomp_b200.simulator import ParallelWriter

    class SlabPull(ParallelWriter):
        def __init__(self, period):
            ParallelWriter.__init__(self, "", period)
            self.calls, self.out = [], None

        def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank_, lastCall):
            self.calls.append((event, step, validRegion, tuple(globalDimensions), rank_, lastCall))
            (o, d) = validRegion
            self.out = grid.saveMember("temp", origin=o, dims=d)

    nz, ny, nx, steps = 12, 5, 6, 4
    data = synth.jacobi_grid(nx, ny, nz, seed=3)
    sim = StripedSimulator(SlabInit(data, steps, 0.75), models.ALL["Jacobi7Cube"], rank=rank, world=world, ghost_width=2,
                           dist=dist, engine=cpu_engine, stream_io=True)   # more than one rank: run() must not stream
    pull = SlabPull(2)
    sim.addWriter(pull)
    sim.run()
    b = slab_bounds(nz, world)
    region = ((0, 0, b[rank]), (nx, ny, b[rank + 1] - b[rank]))
    want_calls = [(0, 0, region, (nx, ny, nz), rank, True), (1, 2, region, (nx, ny, nz), rank, True),
                  (2, 4, region, (nx, ny, nz), rank, True)]
    if sim.streamed_runs != 0:
        failures.append("streamed on %d ranks" % world)
    if pull.calls != want_calls or not np.array_equal(pull.out, oracle_py.jacobi(7, False, data, steps, edge=0.75)[b[rank]:b[rank + 1]]):
        failures.append("parallel writer rank %d: %r" % (rank, pull.calls))

    # advance() calls that are not multiples of the ghost width (a writer period of 5 with ghost width 2; bench.py's
    # --warmup 5): every round stays a fused, overlapped one — no blocking exchange after the very first, no
    # single-sweep launches except the one-sweep tail of a call (round 1's SCALE run lost 10 points to this)
    nz, ny, nx = 24, 5, 6
    data = synth.jacobi_grid(nx, ny, nz, seed=11)
    sim = StripedSimulator(SlabInit(data, 25, 0.75), models.ALL["Jacobi27Cube"], rank=rank, world=world, ghost_width=2,
                           dist=dist, engine=cpu_engine)
    log = []
    dev = sim.grid.dev
    real_box, real_step, real_exchange = dev.update_box, dev.step, sim.halo.exchange
    dev.update_box = lambda *a, **kw: (log.append(("box", kw.get("n_sweeps", 1))), real_box(*a, **kw))[1]
    dev.step = lambda *a, **kw: (log.append(("step", kw.get("n_steps", 1))), real_step(*a, **kw))[1]
    sim.halo.exchange = lambda: (log.append(("exchange", 0)), real_exchange())[1]
    sim.advance(5)
    warm = list(log)
    del log[:]
    sim.advance(20)
    if [e for e in warm if e[0] != "box"] != [("exchange", 0)] or sorted(set(n for _, n in warm if _ == "box")) != [1, 2]:
        failures.append("misaligned advance, warm-up schedule: %r" % (warm,))
    if any(e != ("box", 2) for e in log) or len(log) != 10 * (3 if 0 < rank < world - 1 else 2):
        failures.append("misaligned advance, timed schedule: %r" % (log,))
    b = slab_bounds(nz, world)
    if not np.array_equal(sim.getGrid().saveMember("temp"), oracle_py.jacobi(27, False, data, 25, edge=0.75)[b[rank]:b[rank + 1]]):
        failures.append("misaligned advance: result rank %d" % rank)

    # streamed run on several ranks: ghost zones as wide as the run is long, filled by the Initializer (it is asked for
    # the slab plus its ghost zones), no exchange while the wavefront passes; writers see the rank's own cells only
    class WindowPull(ParallelWriter):
        def __init__(self, shape, member):
            ParallelWriter.__init__(self, "", 1 << 30)
            self.out, self.member, self.regions = np.zeros(shape), member, []

        def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank_, lastCall):
            (o, d) = validRegion
            self.regions.append((event, o[-1], d[-1], lastCall))
            if event == 2:
                z = o[-1] - self.z0
                grid.saveMember(self.member, origin=o, dims=d, out=self.out[z:z + d[-1]])

    for kind, steps, shape, chunks, depth in [(27, 4, (36, 6, 7), 4, 2), (7, 5, (48, 5, 6), 3, 4), (6, 3, (24, 4, 5), 6, 1)]:
        nz, ny, nx = shape
        data = synth.jacobi_grid(nx, ny, nz, seed=kind + 7)
        b = slab_bounds(nz, world)
        sim = StripedSimulator(SlabInit(data, steps, 0.75), models.ALL["Jacobi%dCube" % kind], rank=rank, world=world,
                               ghost_width=steps, dist=dist, engine=cpu_engine, stream_io=True, stream_depth=depth, stream_chunks=chunks)
        pull = WindowPull((b[rank + 1] - b[rank], ny, nx), "temp")
        pull.z0 = b[rank]
        sim.addWriter(pull)
        sim.run()
        want = oracle_py.jacobi(kind, False, data, steps, edge=0.75)[b[rank]:b[rank + 1]]
        done = [r for r in pull.regions if r[0] == 2]
        if sim.streamed_runs != 1 or not np.array_equal(pull.out, want) or not np.array_equal(sim.getGrid().saveMember("temp"), want):
            failures.append("streamed on %d ranks, kind %d rank %d" % (world, kind, rank))
        if min(r[1] for r in pull.regions) < b[rank] or max(r[1] + r[2] for r in pull.regions) > b[rank + 1] or \
                sum(r[2] for r in done) != b[rank + 1] - b[rank] or [r[3] for r in done].count(True) != 1 or not done[-1][3]:
            failures.append("streamed on %d ranks: writer regions %r" % (world, pull.regions))
        # the plain schedule still works on the same simulator afterwards (ghost zones were used up, not left 'valid')
        sim.stream_io = False
        sim.run()
        if not np.array_equal(sim.getGrid().saveMember("temp"), want):
            failures.append("plain run after a streamed one, kind %d rank %d" % (kind, rank))

    # n-body: slabs of BoxCell containers, one ghost plane of containers (counts + particles) per side and sweep;
    # velocities large enough that particles change containers and slabs
    class CellInit(SimpleInitializer):
        def __init__(self, counts, parts, steps):
            SimpleInitializer.__init__(self, counts.shape[::-1], steps)
            self.counts, self.parts = counts, parts

        def grid(self, target):
            o, d = target.boundingBox()
            sl = tuple(slice(o[i], o[i] + d[i]) for i in reversed(range(3)))
            target.loadCells(self.counts[sl], self.parts[sl], origin=o)

    for real, dims, steps, vel in [(np.float32, (4, 3, 7), 6, 10.0), (np.float64, (3, 4, 6), 5, 12.0)]:
        c, p = synth.nbody_cells(*dims, vel=vel, dtype=real)
        model = (models.NBodyF if real == np.float32 else models.NBodyD).with_params(dt=0.01)
        sim = StripedSimulator(CellInit(c, p, steps), model, rank=rank, world=world, ghost_width=1, dist=dist, engine=cpu_engine)
        sim.run()
        gc, gp = sim.getGrid().saveCells()
        wc, wp = oracle_py.nbody(c, p, steps, dt=0.01)
        b = slab_bounds(dims[2], world)
        moved = not np.array_equal(wc, c)
        if not (moved and np.array_equal(gc, wc[b[rank]:b[rank + 1]]) and
                np.array_equal(gp.view(np.uint8), np.ascontiguousarray(wp[b[rank]:b[rank + 1]]).view(np.uint8))):
            failures.append("nbody %s rank %d" % (np.dtype(real).name, rank))

    with open("%s.%d" % (out_path, rank), "w") as f:
        f.write("FAIL " + "; ".join(failures) if failures else "OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
