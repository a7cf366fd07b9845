"""Host-side mirror of the reference's plugin API for the cell-update path.

Same names, argument meaning and event protocol as the reference, so the parity tests read like
the reference's own (parallelization/test/unit/serialsimulatortest.h):

  Initializer / SimpleInitializer   io/initializer.h:22-71, io/simpleinitializer.h:13
  Writer                            io/writer.h:32-88    (period 0 -> ValueError = std::invalid_argument)
  Steerer / SteererFeedback         io/steerer.h:26-121
  B200Grid                          storage/gridbase.h:71-309 (GridBase as plugins see it)
  B200Simulator                     parallelization/simulator.h:28-106, monolithicsimulator.h:17-52,
                                    serialsimulator.h:48-187 (event order, double initialisation)

The authoritative drop-in for C++ users is include/libgeodecomp_b200/b200simulator.h; this module
drives the same C ABI from Python for tests and bench.py. All compute happens in libb200geo.so.
"""
import numpy as np

from . import capi

WRITER_INITIALIZED, WRITER_STEP_FINISHED, WRITER_ALL_DONE = 0, 1, 2
STEERER_INITIALIZED, STEERER_NEXT_STEP, STEERER_ALL_DONE = 0, 1, 2


class Initializer:
    """io/initializer.h:22-71"""

    def grid(self, target):
        raise NotImplementedError

    def gridDimensions(self):
        raise NotImplementedError

    def gridBox(self):
        d = self.gridDimensions()
        return (tuple(0 for _ in d), tuple(d))

    def startStep(self):
        raise NotImplementedError

    def maxSteps(self):
        raise NotImplementedError


class SimpleInitializer(Initializer):
    """io/simpleinitializer.h:13-48"""

    def __init__(self, dimensions, steps=300):
        self.dimensions, self.steps = tuple(int(v) for v in dimensions), int(steps)

    def gridDimensions(self):
        return self.dimensions

    def maxSteps(self):
        return self.steps

    def startStep(self):
        return 0


class Writer:
    """io/writer.h:32-88"""

    def __init__(self, prefix="", period=1):
        if period == 0:
            raise ValueError("period must be positive")
        self.prefix, self.period = prefix, period

    def getPeriod(self):
        return self.period

    def stepFinished(self, grid, step, event):
        raise NotImplementedError


class ParallelWriter(Writer):
    """io/parallelwriter.h:29-107: a writer that is handed the grid piece by piece —
    stepFinished(grid, validRegion, globalDimensions, step, event, rank, lastCall) may be called several times per
    step, each time with another part of the simulation space, `lastCall` on the final one. `grid` is then a
    GridWindow whose boundingBox() is `validRegion` (origin, dims). Reads through a GridWindow are ASYNCHRONOUS
    (stream ordered): the host buffers are complete when the simulator's run() returns."""

    def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank, lastCall):
        raise NotImplementedError

    def stepFinished(self, grid, step, event):
        # handed the whole grid at once (B200Simulator, StripedSimulator without streaming)
        self.stepFinishedRegion(grid, grid.boundingBox(), getattr(grid, "global_dims", grid.dimensions()), step,
                                event, 0, True)


class SteererFeedback:
    """io/steerer.h:45-62"""

    def __init__(self):
        self._ended = False

    def endSimulation(self):
        self._ended = True

    def simulationEnded(self):
        return self._ended


class Steerer:
    """io/steerer.h:26-121"""

    def __init__(self, period=1):
        self.period = period
        self.region = None

    def getPeriod(self):
        return self.period

    def setRegion(self, region):
        self.region = region

    def nextStep(self, grid, validRegion, globalDimensions, step, event, rank, lastCall, feedback):
        raise NotImplementedError


TB_DEPTH = 4  # deepest temporal blocking the Jacobi kernels offer (b200geo_set_tuning "jacobi.tb")


class B200Grid:
    """GridBase<CELL, DIM> as Initializers, Writers and Steerers see it (storage/gridbase.h:71-309),
    backed by the device-resident SoA grid. Coordinates are (x, y[, z]) like Coord<DIM>.
    `origin`/`global_dims` place this grid inside a larger simulation space (slab partitions)."""

    def __init__(self, model, dims, device=0, ghost_z=None, z_modes=None, origin=None, global_dims=None,
                 engine=None):
        self.model = model
        self.dims = tuple(int(v) for v in dims)
        assert len(self.dims) == model.dim
        self.origin = tuple(origin) if origin is not None else tuple(0 for _ in self.dims)
        self.global_dims = tuple(global_dims) if global_dims is not None else self.dims
        wrap = capi.GHOST_WRAP if model.wraps else capi.GHOST_EDGE
        modes = [[wrap, wrap] for _ in range(3)]
        r = model.radius
        ghost = [r, r, r if model.dim == 3 else 0]
        if model.wraps and model.kernel in (capi.KERNEL_JACOBI6, capi.KERNEL_JACOBI7, capi.KERNEL_JACOBI27):
            # periodic images TB_DEPTH cells wide let the temporal-blocked kernel take that many
            # sweeps per launch on a Torus (a wrap ghost may not be wider than the grid itself)
            ghost = [max(r, min(TB_DEPTH, d)) for d in self.dims]
        if model.dim == 2:
            modes[2] = [capi.GHOST_EDGE, capi.GHOST_EDGE]
        last = model.dim - 1
        if ghost_z is not None:
            ghost[last] = int(ghost_z)
        if z_modes is not None:
            modes[last] = list(z_modes)
        self.engine = engine = engine or capi
        self.dev = engine.DeviceGrid(self.dims, model.member_bytes, ghost=ghost, ghost_mode=modes, device=device)
        self.ghost, self.modes = ghost, modes
        self.setEdge(model.default_cell)

    # -- GridBase interface
    def boundingBox(self):
        return (self.origin, self.dims)

    def dimensions(self):
        return self.dims

    def _local3(self, coord):
        c = [int(coord[i]) - self.origin[i] for i in range(self.model.dim)]
        return c + [0] * (3 - len(c))

    def setEdge(self, cell):
        self._edge = self.model.cell_to_bytes(cell)
        self.dev.set_edge(self._edge)

    def getEdge(self):
        return np.frombuffer(self.dev.get_edge(), dtype=self.model.cell_dtype)[0]

    def set(self, coord, cell):
        """set(Coord, CELL) — one cell; like CUDAGrid::set this is a tiny transfer per call
        (storage/cudagrid.h:97-113). Use loadMember / set_streak for bulk initialisation."""
        raw = self.model.cell_to_bytes(cell)
        off = 0
        for m, (_, t) in enumerate(self.model.members):
            v = np.frombuffer(raw[off:off + t.itemsize], dtype=t).copy()
            self.dev.load_member(m, v, self._local3(coord), (1, 1, 1))
            off += t.itemsize

    def set_streak(self, origin, cells):
        """set(Streak, const CELL*): cells is a structured array of the model's cell dtype."""
        cells = np.asarray(cells, dtype=self.model.cell_dtype)
        for m, (n, _) in enumerate(self.model.members):
            self.dev.load_member(m, np.ascontiguousarray(cells[n]), self._local3(origin), (len(cells), 1, 1))

    def get(self, coord):
        out = np.zeros((), dtype=self.model.cell_dtype)
        for m, (n, t) in enumerate(self.model.members):
            v = np.zeros(1, dtype=t)
            self.dev.save_member(m, v, self._local3(coord), (1, 1, 1))
            self.engine.sync()
            out[n] = v[0]
        return out

    def get_streak(self, origin, length):
        out = np.zeros(length, dtype=self.model.cell_dtype)
        for m, (n, t) in enumerate(self.model.members):
            v = np.zeros(length, dtype=t)
            self.dev.save_member(m, v, self._local3(origin), (length, 1, 1))
            self.engine.sync()
            out[n] = v
        return out

    def _box(self, origin, dims):
        if origin is None:
            origin = self.origin
        o = self._local3(origin)
        if dims is None:
            dims = [self.dims[i] - o[i] for i in range(self.model.dim)]
        d = list(dims) + [1] * (3 - len(dims))
        return o, d

    def loadMember(self, name, array, origin=None, location=capi.HOST):
        """GridBase::loadMember for a box region: dense array, x fastest ([z][y][x])."""
        m = self.model.member_index(name)
        t = self.model.members[m][1]
        if location == capi.HOST:
            array = np.ascontiguousarray(array, dtype=t)
            dims = array.shape[::-1]
        else:
            dims = tuple(array.shape)[::-1]
        o, d = self._box(origin, dims)
        self.dev.load_member(m, array, o, d, location=location)

    def saveMember(self, name, origin=None, dims=None, out=None, location=capi.HOST):
        m = self.model.member_index(name)
        t = self.model.members[m][1]
        o, d = self._box(origin, dims)
        if out is None:
            out = np.empty(tuple(d[:self.model.dim])[::-1], dtype=t)
        self.dev.save_member(m, out, o, d, location=location)
        self.engine.sync()
        return out

    def _streaks3(self, streaks):
        st = np.asarray(streaks, dtype=np.int64)
        if self.model.dim == 2:   # {x, y, endX} -> {x, y, 0, endX}
            st = st.reshape(-1, 3)
            st = np.stack([st[:, 0], st[:, 1], np.zeros(len(st), dtype=np.int64), st[:, 2]], axis=1)
        st = st.reshape(-1, 4).copy()
        st[:, 0] -= self.origin[0]
        st[:, 3] -= self.origin[0]
        st[:, 1] -= self.origin[1]
        if self.model.dim == 3:
            st[:, 2] -= self.origin[2]
        return st.astype(np.int32)

    def saveRegion(self, streaks):
        """GridBase::saveRegion(std::vector<char>*, Region): member-major bytes (soagrid.h:523-547)."""
        st = self._streaks3(streaks)
        n = int((st[:, 3] - st[:, 0]).sum())
        buf = np.zeros(n * self.model.cell_dtype.itemsize, dtype=np.uint8)
        self.dev.save_region(st, buf)
        self.engine.sync()
        return buf

    def loadRegion(self, buf, streaks):
        st = self._streaks3(streaks)
        n = int((st[:, 3] - st[:, 0]).sum())
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        if buf.size != n * self.model.cell_dtype.itemsize:
            raise ValueError("buffer size does not match region")
        self.dev.load_region(st, buf)
        self.engine.sync()

    def to_raw(self):
        """whole interior as the member-major byte stream (one dense array per member)."""
        return [self.saveMember(n) for n, _ in self.model.members]


class GridWindow:
    """The planes [z0, z1) (rows in 2-D) of a B200Grid along its last axis, as the GridBase an Initializer or a
    ParallelWriter is handed for a sub-box: Initializer::grid 'may be called for sub-boxes and must only touch cells
    inside boundingBox()' (io/initializer.h:38-71), ParallelWriter::stepFinished gets a validRegion
    (io/parallelwriter.h:92-99). Transfers are enqueued on `stream` and touch ONE buffer (the grid's current one at
    the time of the call) — the building block of StripedSimulator's streamed run, where host<->device copies of
    one part of the space overlap the sweeps over another. Members the kernel never rewrites
    (model.invariant_members, e.g. the LBM cell's `state`) go into BOTH buffers, device to device."""

    def __init__(self, grid, z0, z1, stream=None):
        self.grid, self.model, self.stream = grid, grid.model, stream
        last = grid.model.dim - 1
        self.origin = tuple(o + (z0 if i == last else 0) for i, o in enumerate(grid.origin))
        self.dims = tuple((z1 - z0) if i == last else d for i, d in enumerate(grid.dims))
        self.global_dims = grid.global_dims

    def boundingBox(self):
        return (self.origin, self.dims)

    def dimensions(self):
        return self.dims

    def setEdge(self, cell):
        # the edge cell belongs to the whole grid; re-stating it per window is a no-op
        if self.model.cell_to_bytes(cell) != self.grid._edge:
            self.grid.engine.sync()
            self.grid.setEdge(cell)
            self.grid.engine.sync()

    def getEdge(self):
        return self.grid.getEdge()

    def _inside(self, origin, dims):
        if origin is None:
            origin = self.origin
        if dims is None:
            dims = [self.dims[i] - (origin[i] - self.origin[i]) for i in range(self.model.dim)]
        for i in range(self.model.dim):
            if origin[i] < self.origin[i] or origin[i] + dims[i] > self.origin[i] + self.dims[i]:
                raise ValueError("box outside the window's boundingBox()")
        return origin, dims

    def loadMember(self, name, array, origin=None, location=capi.HOST):
        m = self.model.member_index(name)
        t = self.model.members[m][1]
        if location == capi.HOST:
            array = np.ascontiguousarray(array, dtype=t)
        origin, dims = self._inside(origin, tuple(array.shape)[::-1])
        o, d = self.grid._box(origin, dims)
        self.grid.dev.load_member(m, array, o, d, location=location, both=m in self.model.invariant_members,
                                  stream=self.stream)

    def set_streak(self, origin, cells):
        """set(Streak, const CELL*): cells is a structured array of the model's cell dtype"""
        cells = np.asarray(cells, dtype=self.model.cell_dtype)
        origin, _ = self._inside(tuple(origin), (len(cells),) + (1,) * (self.model.dim - 1))
        for m, (n, _) in enumerate(self.model.members):
            self.grid.dev.load_member(m, np.ascontiguousarray(cells[n]), self.grid._local3(origin), (len(cells), 1, 1),
                                      both=m in self.model.invariant_members, stream=self.stream)

    def set(self, coord, cell):
        one = np.zeros(1, dtype=self.model.cell_dtype)
        one[0] = np.frombuffer(self.model.cell_to_bytes(cell), dtype=self.model.cell_dtype)[0]
        self.set_streak(coord, one)

    def get_streak(self, origin, length):
        """get(Streak, CELL*): synchronous — waits for the window's stream"""
        origin, _ = self._inside(tuple(origin), (length,) + (1,) * (self.model.dim - 1))
        out = np.zeros(length, dtype=self.model.cell_dtype)
        for m, (n, t) in enumerate(self.model.members):
            v = np.zeros(length, dtype=t)
            self.grid.dev.save_member(m, v, self.grid._local3(origin), (length, 1, 1), stream=self.stream)
            self.grid.engine.sync(self.stream)
            out[n] = v
        return out

    def get(self, coord):
        return self.get_streak(coord, 1)[0]

    def saveMember(self, name, origin=None, dims=None, out=None, location=capi.HOST):
        m = self.model.member_index(name)
        origin, dims = self._inside(origin, dims)
        o, d = self.grid._box(origin, dims)
        if out is None:
            raise ValueError("GridWindow.saveMember is asynchronous: pass the (pinned) array to fill as out=")
        self.grid.dev.save_member(m, out, o, d, location=location, stream=self.stream)
        return out


class B200Simulator:
    """MonolithicSimulator on one B200: same observable behaviour as SerialSimulator
    (parallelization/serialsimulator.h:48-187)."""

    def __init__(self, initializer, model, device=0, engine=None):
        self.initializer, self.model = initializer, model
        self.NANO_STEPS = model.nano_steps
        dims = initializer.gridDimensions()
        # engine: tests hand in a stand-in for the C ABI (tests/cpu_engine.py); the product always runs on capi
        self.grid = model.grid_class(model, dims, device=device, **({"engine": engine} if engine is not None else {}))
        self._sync = capi.sync if engine is None else getattr(engine, "sync", lambda: None)
        # SerialSimulator initialises both grids (serialsimulator.h:54-57); loads write both buffers
        initializer.grid(self.grid)
        self.stepNum = initializer.startStep()
        self.writers, self.steerers = [], []
        self.grid.dev.stats_enable(False)

    # -- Simulator interface
    def addWriter(self, writer):
        self.writers.append(writer)

    def addSteerer(self, steerer):
        self.steerers.append(steerer)

    def getStep(self):
        return self.stepNum

    def getGrid(self):
        self._sync()
        return self.grid

    def gatherStatistics(self):
        return [self.grid.dev.stats()]

    def step(self, feedback=None):
        feedback = feedback or SteererFeedback()
        self._handleInput(STEERER_NEXT_STEP, feedback)
        self._advance(1)
        self._afterStep()

    def run(self):
        self.initializer.grid(self.grid)
        self.stepNum = self.initializer.startStep()
        for s in self.steerers:
            s.setRegion(self.grid.boundingBox())
        feedback = SteererFeedback()
        self._handleInput(STEERER_INITIALIZED, feedback)
        self._handleOutput(WRITER_INITIALIZED)
        maxSteps = self.initializer.maxSteps()
        while self.stepNum < maxSteps:
            if feedback.simulationEnded():
                break
            self._handleInput(STEERER_NEXT_STEP, feedback)
            # fuse all steps up to the next observable event into one engine call
            n = self._steps_to_next_event(maxSteps)
            self._advance(n)
            self._afterStep()
        self._handleInput(STEERER_ALL_DONE, feedback)
        self._sync()

    # -- internals
    def _steps_to_next_event(self, maxSteps):
        n = maxSteps - self.stepNum
        for w in self.writers:
            n = min(n, w.getPeriod() - self.stepNum % w.getPeriod())
        for s in self.steerers:
            n = min(n, s.getPeriod() - self.stepNum % s.getPeriod())
        return max(1, n)

    def _advance(self, steps):
        self.grid.dev.step(self.model.kernel, n_steps=steps * self.NANO_STEPS, first_nano_step=0,
                           params=self.model.step_params(True))
        self.stepNum += steps

    def _afterStep(self):
        event = WRITER_ALL_DONE if self.stepNum == self.initializer.maxSteps() else WRITER_STEP_FINISHED
        self._handleOutput(event)

    def _handleOutput(self, event):
        for w in self.writers:
            if event != WRITER_STEP_FINISHED or self.stepNum % w.getPeriod() == 0:
                w.stepFinished(self.getGrid(), self.stepNum, event)

    def _handleInput(self, event, feedback):
        for s in self.steerers:
            if event != STEERER_NEXT_STEP or self.stepNum % s.getPeriod() == 0:
                s.nextStep(self.grid, self.grid.boundingBox(), self.grid.dims, self.stepNum, event, 0, True, feedback)
