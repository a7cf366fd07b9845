"""GridBase as plugins see it for grids of BoxCell containers (short-range n-body), backed by the
device-resident cell-list grid (csrc/nbody.cu).

Mirrors Grid<BoxCell<FixedArray<Particle, N> > > behind GridBase<CELL, 3> (storage/gridbase.h:71-309,
storage/boxcell.h:21-178): set/get carry whole containers — here a container is an (n, 6) array
of particles (pos x,y,z, vel x,y,z); origin and dimension of a container follow from its coordinate
(origin = coord * cell_edge), which is how the bound model's Initializer constructs them
(oracle/models/nbody.h). Bulk variants move boxes of containers in the interchange format of
include/b200geo.h (counts [z][y][x], particles [z][y][x][capacity][6]).
"""
import numpy as np

from . import capi


class BoxGrid:
    def __init__(self, model, dims, device=0, ghost_z=None, z_modes=None, origin=None, global_dims=None, engine=None):
        self.model = model
        self.dims = tuple(int(v) for v in dims)
        assert len(self.dims) == 3
        self.origin = tuple(origin) if origin is not None else (0, 0, 0)
        self.global_dims = tuple(global_dims) if global_dims is not None else self.dims
        if ghost_z not in (None, 1):
            raise ValueError("BoxCell grids have a ghost zone of one container")
        modes = [[capi.GHOST_EDGE, capi.GHOST_EDGE] for _ in range(3)]
        if z_modes is not None:
            modes[2] = list(z_modes)
        self.engine = engine or capi
        self.dev = self.engine.DeviceBoxGrid(self.dims, model.capacity, model.real.itemsize, model.cell_edge,
                                             cell_origin=self.origin, ghost_mode=modes, device=device)
        self.ghost, self.modes = [1, 1, 1], modes

    # -- GridBase interface
    def boundingBox(self):
        return (self.origin, self.dims)

    def dimensions(self):
        return self.dims

    def _local(self, coord):
        return [int(coord[i]) - self.origin[i] for i in range(3)]

    def setEdge(self, cell):
        if len(cell):
            raise capi.LogicError("the edge container of a BoxCell grid is empty")

    def getEdge(self):
        return np.zeros((0, 6), dtype=self.model.real)

    def set(self, coord, cell):
        """set(Coord, BoxCell): cell = (n, 6) particles"""
        cell = np.asarray(cell, dtype=self.model.real).reshape(-1, 6)
        if len(cell) > self.model.capacity:
            raise IndexError("capacity exceeded")
        parts = np.zeros((1, 1, 1, self.model.capacity, 6), dtype=self.model.real)
        parts[0, 0, 0, :len(cell)] = cell
        self.dev.load(np.array([len(cell)], dtype=np.int32), parts, self._local(coord), (1, 1, 1))

    def get(self, coord):
        c, p = self.saveCells(coord, (1, 1, 1))
        return p[0, 0, 0, :c[0, 0, 0]].copy()

    def loadCells(self, counts, particles, origin=None):
        """containers of a box: counts [dz][dy][dx] int32, particles [dz][dy][dx][capacity][6]"""
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        particles = np.ascontiguousarray(particles, dtype=self.model.real)
        assert particles.shape == counts.shape + (self.model.capacity, 6)
        if counts.size and counts.max() > self.model.capacity:
            raise IndexError("capacity exceeded")
        o = self._local(origin if origin is not None else self.origin)
        self.dev.load(counts, particles, o, counts.shape[::-1])

    def saveCells(self, origin=None, dims=None, out=None):
        o = self._local(origin if origin is not None else self.origin)
        d = tuple(dims) if dims is not None else tuple(self.dims[i] - o[i] for i in range(3))
        self.dev.check()
        if out is None:
            out = (np.empty(d[::-1], dtype=np.int32), np.empty(d[::-1] + (self.model.capacity, 6), dtype=self.model.real))
        self.dev.save(out[0], out[1], o, d)
        return out
