"""StripedSimulator's streamed run (libgeodecomp_b200/striping.py::_run_streamed): Initializer -> sweeps -> ParallelWriter
pipelined chunk by chunk along the last axis with a time-skewed schedule. Here on the CPU engine (tests/cpu_engine.py:
the DeviceGrid interface on numpy arrays, update_box = the oracle on the box plus its halo, current -> scratch): the
schedule — which planes are swept when, out of which buffer — must give the oracle's result bit for bit, for every
fusing depth, remainder level, chunk size and step count, and the plugins must see the windows they are promised."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import cpu_engine  # noqa: E402
from libgeodecomp_b200 import models, synth  # noqa: E402
from libgeodecomp_b200.simulator import ParallelWriter, SimpleInitializer, Steerer, Writer  # noqa: E402
from libgeodecomp_b200.striping import StripedSimulator  # noqa: E402
from oracle import oracle_py  # noqa: E402


class BoxInit(SimpleInitializer):
    """initialises only the cells inside target.boundingBox() (io/initializer.h:38-44)"""

    def __init__(self, data, steps, edge):
        SimpleInitializer.__init__(self, data.shape[::-1], steps)
        self.data, self.edge, self.boxes = data, edge, []

    def grid(self, target):
        (ox, oy, oz), (dx, dy, dz) = target.boundingBox()
        self.boxes.append((oz, dz))
        target.setEdge(self.edge)
        target.loadMember("temp", self.data[oz:oz + dz, oy:oy + dy, ox:ox + dx], origin=(ox, oy, oz))


class Pull(ParallelWriter):
    """collects the final grid window by window"""

    def __init__(self, shape, period):
        ParallelWriter.__init__(self, "", period)
        self.out = np.full(shape, np.nan)
        self.calls = []

    def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank, lastCall):
        (ox, oy, oz), (dx, dy, dz) = validRegion
        assert grid.boundingBox() == validRegion
        self.calls.append((event, step, oz, dz, lastCall))
        if event == 2:
            grid.saveMember("temp", origin=(ox, oy, oz), dims=(dx, dy, dz), out=self.out[oz:oz + dz])


@pytest.mark.parametrize("kind,depth,steps,shape,chunks", [
    (7, 1, 5, (40, 6, 9), 5),
    (7, 4, 10, (64, 5, 8), 4),       # 2 full levels of 4 + a remainder level of 2 (goes first)
    (7, 4, 3, (48, 5, 8), 3),        # fewer sweeps than the fusing depth
    (27, 2, 7, (36, 7, 6), 6),       # odd number of sweeps: levels [1, 2, 2, 2]
    (27, 2, 8, (33, 7, 6), 4),       # ragged last chunk
    (6, 3, 9, (50, 4, 7), 16),       # chunks shorter than asked for: clamped to 2 * depth planes
    (27, 2, 40, (30, 5, 5), 3),      # the skew (40 planes) is longer than the grid
])
def test_streamed_run_equals_the_oracle(kind, depth, steps, shape, chunks):
    nz, ny, nx = shape
    data = synth.jacobi_grid(nx, ny, nz, seed=kind)
    model = models.ALL["Jacobi%dCube" % kind]
    init = BoxInit(data, steps, 0.75)
    sim = StripedSimulator(init, model, engine=cpu_engine, stream_io=True, stream_depth=depth, stream_chunks=chunks)
    init.boxes.clear()
    pull = Pull(shape, steps)
    sim.addWriter(pull)
    sim.run()
    assert sim.streamed_runs == 1 and sim.getStep() == steps
    want = oracle_py.jacobi(kind, False, data, steps, edge=0.75)
    assert np.array_equal(pull.out, want)
    # the device grid ends up with the final state in its current buffer, as after a plain run
    assert np.array_equal(sim.getGrid().saveMember("temp"), want)
    # the Initializer saw disjoint windows covering the axis once, in order
    assert [b[0] for b in init.boxes] == list(np.cumsum([0] + [b[1] for b in init.boxes[:-1]]))
    assert sum(b[1] for b in init.boxes) == nz and len(init.boxes) > 1
    # the writer saw WRITER_INITIALIZED per uploaded window and WRITER_ALL_DONE per finished one, lastCall exactly once each
    for event, step in ((0, 0), (2, steps)):
        calls = [c for c in pull.calls if c[0] == event]
        assert all(c[1] == step for c in calls)
        assert sum(c[3] for c in calls) == nz and [c[4] for c in calls].count(True) == 1 and calls[-1][4]
    # and a second run() re-initialises and gives the same result (serialsimulator.h:100)
    pull.out[...] = np.nan
    sim.run()
    assert sim.streamed_runs == 2 and np.array_equal(pull.out, want)


def test_streamed_lbm_equals_the_oracle():
    nx, ny, nz, steps = 8, 7, 24, 5
    raw = synth.lbm_grid(nx, ny, nz, noise=0.01)

    class LBMInit(SimpleInitializer):
        def grid(self, target):
            (ox, oy, oz), (dx, dy, dz) = target.boundingBox()
            for m, (name, t) in enumerate(models.LBMCellF.members):
                target.loadMember(name, raw[m, oz:oz + dz, oy:oy + dy, ox:ox + dx].view(t), origin=(ox, oy, oz))

    out = np.zeros_like(raw)

    class LBMPull(ParallelWriter):
        def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank, lastCall):
            (ox, oy, oz), (dx, dy, dz) = validRegion
            if event == 2:
                for m, (name, t) in enumerate(models.LBMCellF.members):
                    grid.saveMember(name, origin=(ox, oy, oz), dims=(dx, dy, dz), out=out[m, oz:oz + dz].view(t))

    sim = StripedSimulator(LBMInit((nx, ny, nz), steps), models.LBMCellF, engine=cpu_engine, stream_io=True, stream_chunks=4)
    sim.addWriter(LBMPull("", steps))
    sim.run()
    assert sim.streamed_runs == 1
    want = oracle_py.lbm(raw, steps)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_streamed_lbm_with_walls_that_changed_after_construction():
    """The LBM kernel never rewrites `state` (nor the walls' density / velocity): a streamed upload has to put them into
    BOTH buffers, or odd levels read what the constructor-time Initializer::grid call left there."""
    nx, ny, nz, steps = 8, 7, 24, 5
    raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
    src = raw.copy()
    state = models.LBMCellF.member_index("state")
    src[state] = 0                       # at construction: no walls at all

    class LBMInit(SimpleInitializer):
        def grid(self, target):
            (ox, oy, oz), (dx, dy, dz) = target.boundingBox()
            for m, (name, t) in enumerate(models.LBMCellF.members):
                target.loadMember(name, src[m, oz:oz + dz, oy:oy + dy, ox:ox + dx].view(t), origin=(ox, oy, oz))

    out = np.zeros_like(raw)

    class LBMPull(ParallelWriter):
        def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank, lastCall):
            (ox, oy, oz), (dx, dy, dz) = validRegion
            if event == 2:
                for m, (name, t) in enumerate(models.LBMCellF.members):
                    grid.saveMember(name, origin=(ox, oy, oz), dims=(dx, dy, dz), out=out[m, oz:oz + dz].view(t))

    sim = StripedSimulator(LBMInit((nx, ny, nz), steps), models.LBMCellF, engine=cpu_engine, stream_io=True, stream_chunks=4)
    sim.addWriter(LBMPull("", steps))
    src[...] = raw                       # the cavity's walls appear before run()
    sim.run()
    assert sim.streamed_runs == 1
    assert np.array_equal(out.view(np.uint32), oracle_py.lbm(raw, steps).view(np.uint32))


def test_a_writer_period_that_falls_due_inside_the_run_is_not_streamed():
    """startStep 5, maxSteps 12, period 10: step 10 is due mid-run (the plain schedule fires WRITER_STEP_FINISHED there)"""
    nz, ny, nx = 24, 5, 6
    data = synth.jacobi_grid(nx, ny, nz)

    class Late(BoxInit):
        def startStep(self):
            return 5

        def maxSteps(self):
            return 12

    for period, streams in ((10, False), (12, True), (13, True), (7, False), (6, False)):
        sim = StripedSimulator(Late(data, 12, 0.75), models.ALL["Jacobi7Cube"], engine=cpu_engine, stream_io=True, stream_chunks=4)
        pull = Pull((nz, ny, nx), period)
        sim.addWriter(pull)
        sim.run()
        assert sim.streamed_runs == (1 if streams else 0), period
        assert np.array_equal(pull.out, oracle_py.jacobi(7, False, data, 7, edge=0.75)), period


def test_run_falls_back_to_the_plain_schedule_when_it_cannot_stream():
    nz, ny, nx, steps = 24, 5, 6, 4
    data = synth.jacobi_grid(nx, ny, nz)
    want = oracle_py.jacobi(7, False, data, steps, edge=0.75)

    class Plain(Writer):
        def stepFinished(self, grid, step, event):
            pass

    cases = {
        "a serial Writer wants the whole grid": dict(writers=[Plain("", 1)]),
        "a ParallelWriter with a period inside the run": dict(writers=[Pull((nz, ny, nx), 2)]),
        "a Steerer may change cells between steps": dict(steerers=[type("S", (Steerer,), {"nextStep": lambda *a: None})(1)]),
        "not asked for": dict(stream_io=False),
        "grid too short for two chunks": dict(stream_chunks=1),
    }
    for why, kw in cases.items():
        sim = StripedSimulator(BoxInit(data, steps, 0.75), models.ALL["Jacobi7Cube"], engine=cpu_engine,
                               stream_io=kw.get("stream_io", True), stream_chunks=kw.get("stream_chunks", 4))
        for w in kw.get("writers", []):
            sim.addWriter(w)
        for s in kw.get("steerers", []):
            sim.addSteerer(s)
        sim.run()
        assert sim.streamed_runs == 0, why
        assert np.array_equal(sim.getGrid().saveMember("temp"), want), why
    # a Torus cannot be streamed either
    sim = StripedSimulator(BoxInit(data, steps, 0.0), models.ALL["Jacobi7Torus"], engine=cpu_engine, stream_io=True)
    sim.run()
    assert sim.streamed_runs == 0
    assert np.array_equal(sim.getGrid().saveMember("temp"), oracle_py.jacobi(7, True, data, steps))


def test_window_rejects_writes_outside_its_bounding_box():
    data = synth.jacobi_grid(6, 5, 24)

    class Greedy(SimpleInitializer):
        def grid(self, target):
            target.loadMember("temp", data, origin=(0, 0, 0))   # the whole grid, whatever the box

    sim = StripedSimulator(Greedy((6, 5, 24), 2), models.ALL["Jacobi7Cube"], engine=cpu_engine, stream_io=True, stream_chunks=4)
    sim.addWriter(Pull((24, 5, 6), 2))
    with pytest.raises(ValueError):
        sim.run()


def test_bench_plugins_work_on_whole_grids_and_on_windows():
    """bench.py's e2e Initializer / Writer pair (host arrays in, host arrays out) through the plain and the streamed
    run, in place (input and output share the arrays, as in bench.py), and on a rank's slab of a larger space."""
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    nz, ny, nx, steps = 40, 5, 6, 6
    data = synth.jacobi_grid(nx, ny, nz)
    want = oracle_py.jacobi(27, False, data, steps)
    for stream_io in (False, True):
        host = {"temp": data.copy()}
        Init, PullWriter = bench.make_plugins(host, host, 2, 0)
        sim = StripedSimulator(Init((nx, ny, nz), steps), models.ALL["Jacobi27Cube"], engine=cpu_engine, stream_io=stream_io,
                               stream_chunks=4)
        sim.writers = [PullWriter("", 1 << 30)]
        sim.run()
        assert sim.streamed_runs == (1 if stream_io else 0)
        assert np.array_equal(host["temp"], want)
    # z0 != 0: the arrays hold the rank's slab only (boundingBox() is in global coordinates)
    host = {"temp": data[10:30].copy()}
    Init, PullWriter = bench.make_plugins(host, host, 2, 10)

    class Target:
        def boundingBox(self):
            return ((0, 0, 14), (nx, ny, 8))

        def loadMember(self, name, array, origin=None):
            assert origin == (0, 0, 14) and np.array_equal(array, data[14:22])

    Init((nx, ny, nz), steps).grid(Target())


def test_initializers_that_write_rows_and_single_cells_into_a_window():
    nz, ny, nx, steps = 24, 4, 5, 3
    data = synth.jacobi_grid(nx, ny, nz)
    model = models.ALL["Jacobi7Cube"]

    class RowInit(SimpleInitializer):
        def grid(self, target):
            (ox, oy, oz), (dx, dy, dz) = target.boundingBox()
            for z in range(oz, oz + dz):
                for y in range(oy, oy + dy):
                    row = np.zeros(dx, dtype=model.cell_dtype)
                    row["temp"] = data[z, y, ox:ox + dx]
                    target.set_streak((ox, y, z), row[:-1])
                    target.set((ox + dx - 1, y, z), float(data[z, y, ox + dx - 1]))

    sim = StripedSimulator(RowInit((nx, ny, nz), steps), model, engine=cpu_engine, stream_io=True, stream_chunks=4)
    pull = Pull((nz, ny, nx), steps)
    sim.addWriter(pull)

    class CellReader(ParallelWriter):
        """reads single cells and rows of the windows it is handed"""
        def __init__(self):
            ParallelWriter.__init__(self, "", steps)
            self.seen = np.full((nz, ny, nx), np.nan)

        def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank, lastCall):
            (ox, oy, oz), (dx, dy, dz) = validRegion
            if event == 2:
                for z in range(oz, oz + dz):
                    self.seen[z, 0, :] = grid.get_streak((ox, oy, z), dx)["temp"]
                    self.seen[z, 1, 2] = grid.get((ox + 2, oy + 1, z))["temp"]
                with pytest.raises(ValueError):
                    grid.get((ox, oy, oz + dz))     # outside the window

    reader = CellReader()
    sim.addWriter(reader)
    sim.run()
    assert sim.streamed_runs == 1
    want = oracle_py.jacobi(7, False, data, steps)
    assert np.array_equal(pull.out, want)
    assert np.array_equal(reader.seen[:, 0, :], want[:, 0, :]) and np.array_equal(reader.seen[:, 1, 2], want[:, 1, 2])


def test_streamed_schedule_random_configurations():
    """40 seeded random combinations of kernel, fusing depth, step count, grid height and chunk count"""
    rng = np.random.default_rng(20260117)
    streamed = 0
    for _ in range(40):
        kind = int(rng.choice([6, 7, 27]))
        depth = int(rng.integers(1, 5))
        steps = int(rng.integers(1, 14))
        nz, ny, nx = int(rng.integers(8, 60)), int(rng.integers(3, 7)), int(rng.integers(3, 8))
        chunks = int(rng.integers(2, 12))
        data = synth.jacobi_grid(nx, ny, nz, seed=int(rng.integers(1, 1000)))
        sim = StripedSimulator(BoxInit(data, steps, 0.75), models.ALL["Jacobi%dCube" % kind], engine=cpu_engine,
                               stream_io=True, stream_depth=depth, stream_chunks=chunks)
        pull = Pull((nz, ny, nx), steps)
        sim.addWriter(pull)
        sim.run()
        want = oracle_py.jacobi(kind, False, data, steps, edge=0.75)
        got = pull.out if sim.streamed_runs else sim.getGrid().saveMember("temp")
        assert np.array_equal(got, want), (kind, depth, steps, (nz, ny, nx), chunks, sim.streamed_runs)
        streamed += sim.streamed_runs
    assert streamed >= 25     # most of them can be streamed; the rest are too short for two chunks and fall back
