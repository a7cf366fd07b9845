"""GridBase as plugins see it for grids of ContainerCell containers (ID-keyed cargo: meshfree / unstructured
models), backed by the device-resident container grid of csrc/container.cu.

Mirrors Grid<ContainerCell<CARGO, SIZE> > behind GridBase<CELL, DIM> (storage/gridbase.h:71-309,
storage/containercell.h:24-218) for the bound cargo model (models.ContainerModel: the mesh element of
src/examples/voronoi/main.cpp:9-118). set/get carry whole containers — here a container is a list of cargo
dicts {"id", "temperature", "influx", "neighborIDs"}, kept in ascending id order like ContainerCell::insert does
(containercell.h:62-88: an id that is already there replaces its cargo). Bulk variants move boxes of containers in
the interchange format of include/b200geo.h.
"""
import numpy as np

from . import capi

FIELDS = capi.ContainerBox.FIELDS


def empty_box(cells_shape, capacity, max_neighbors):
    """the six arrays of a box of empty containers; cells_shape = [dz][dy][dx] (or [dy][dx])"""
    shape = tuple(cells_shape)
    return {"counts": np.zeros(shape, dtype=np.int32),
            "ids": np.zeros(shape + (capacity,), dtype=np.int32),
            "values": np.zeros(shape + (capacity,), dtype=np.float64),
            "influx": np.zeros(shape + (capacity,), dtype=np.float64),
            "nb_counts": np.zeros(shape + (capacity,), dtype=np.int32),
            "nb_ids": np.zeros(shape + (capacity, max_neighbors), dtype=np.int32)}


class ContainerGrid:
    def __init__(self, model, dims, device=0, engine=None, **_):
        self.model = model
        dims = tuple(int(v) for v in dims)
        if len(dims) != model.dim:
            raise ValueError("a %d-D cargo model on a %d-D grid" % (model.dim, len(dims)))
        self.dims = dims
        self.dims3 = dims + (1,) * (3 - len(dims))
        self.origin = (0,) * len(dims)
        mode = capi.GHOST_WRAP if model.wraps else capi.GHOST_EDGE
        modes = [[mode, mode] if i < model.dim else [capi.GHOST_EDGE, capi.GHOST_EDGE] for i in range(3)]
        self.engine = engine or capi
        self.dev = self.engine.DeviceContainerGrid(self.dims3, model.capacity, model.max_neighbors, n_dims=model.dim,
                                                   ghost_mode=modes, device=device)

    # -- GridBase interface
    def boundingBox(self):
        return (self.origin, self.dims)

    def dimensions(self):
        return self.dims

    def _pad(self, v, fill):
        v = tuple(int(x) for x in v)
        return v + (fill,) * (3 - len(v))

    def _pack(self, cell):
        """one container (list of cargo dicts) -> the six arrays of a 1 x 1 x 1 box; insert() semantics"""
        cap, nb = self.model.capacity, self.model.max_neighbors
        by_id = {}
        for cargo in cell:
            by_id[int(cargo["id"])] = cargo
        if len(by_id) > cap:
            raise capi.LogicError("ContainerCell capacity exeeded")
        box = empty_box((1, 1, 1), cap, nb)
        box["counts"][0, 0, 0] = len(by_id)
        for s, key in enumerate(sorted(by_id)):
            cargo = by_id[key]
            n = list(cargo.get("neighborIDs", ()))
            if len(n) > nb:
                raise IndexError("FixedArray capacity exceeded")
            box["ids"][0, 0, 0, s] = key
            box["values"][0, 0, 0, s] = cargo.get("temperature", 0.0)
            box["influx"][0, 0, 0, s] = cargo.get("influx", 0.0)
            box["nb_counts"][0, 0, 0, s] = len(n)
            box["nb_ids"][0, 0, 0, s, :len(n)] = n
        return box

    @staticmethod
    def _unpack(box):
        n = int(box["counts"].reshape(-1)[0])
        ids, val, inf = box["ids"].reshape(-1), box["values"].reshape(-1), box["influx"].reshape(-1)
        nbc, nbi = box["nb_counts"].reshape(-1), box["nb_ids"].reshape(len(ids), -1)
        return [{"id": int(ids[s]), "temperature": float(val[s]), "influx": float(inf[s]),
                 "neighborIDs": [int(v) for v in nbi[s, :nbc[s]]]} for s in range(n)]

    def setEdge(self, cell):
        self.dev.set_edge(self._pack(cell))

    def getEdge(self):
        box = empty_box((1, 1, 1), self.model.capacity, self.model.max_neighbors)
        self.dev.get_edge(box)
        return self._unpack(box)

    def set(self, coord, cell):
        self.dev.load(self._pack(cell), self._pad(coord, 0), (1, 1, 1))

    def get(self, coord):
        box = empty_box((1, 1, 1), self.model.capacity, self.model.max_neighbors)
        self.dev.save(box, self._pad(coord, 0), (1, 1, 1))
        return self._unpack(box)

    def loadCells(self, box, origin=None):
        """containers of a box in the interchange format: dict of the six arrays, cells in [dz][dy][dx] order"""
        counts = np.ascontiguousarray(box["counts"], dtype=np.int32)
        shape = counts.shape
        cap, nb = self.model.capacity, self.model.max_neighbors
        want = {"counts": (shape, np.int32), "ids": (shape + (cap,), np.int32), "values": (shape + (cap,), np.float64),
                "influx": (shape + (cap,), np.float64), "nb_counts": (shape + (cap,), np.int32),
                "nb_ids": (shape + (cap, nb), np.int32)}
        arrays = {}
        for name in FIELDS:
            a = np.ascontiguousarray(box[name], dtype=want[name][1])
            if a.shape != want[name][0]:
                raise ValueError("%s has shape %r, expected %r" % (name, a.shape, want[name][0]))
            arrays[name] = a
        if counts.size and counts.max() > cap:
            raise capi.LogicError("ContainerCell capacity exeeded")
        d = self._pad(shape[::-1], 1)
        self.dev.load(arrays, self._pad(origin if origin is not None else self.origin, 0), d)

    def saveCells(self, origin=None, dims=None, fields=FIELDS):
        """the containers of a box; `fields` limits the arrays fetched (a Writer that wants the temperatures asks for
        ("counts", "values"))"""
        o = self._pad(origin if origin is not None else self.origin, 0)
        d = self._pad(dims, 1) if dims is not None else tuple(self.dims3[i] - o[i] for i in range(3))
        shape = tuple(d[:len(self.dims)][::-1])
        box = empty_box(shape, self.model.capacity, self.model.max_neighbors)
        self.dev.save({n: (box[n] if n in fields else None) for n in FIELDS}, o, d)
        return {n: box[n] for n in FIELDS if n in fields}
