"""Model bindings: the Python twin of B200KernelBinding<CELL> (include/libgeodecomp_b200/).

A binding states what the reference learns from Cell::API (misc/apitraits.h:191-1168) and from
LIBFLATARRAY_REGISTER_SOA — dimensionality, topology, stencil radius, nano steps, the SoA member
table in registration order — plus which hand-written kernel family implements the cell's
update()/updateLineX(). The cell classes themselves live in oracle/models/*.h.
"""
import numpy as np

from . import capi


class CellModel:
    def __init__(self, name, dim, topology, members, kernel, edge=None, nano_steps=1, radius=1, ref_model=None):
        self.name, self.dim, self.topology = name, dim, topology
        self.members = [(n, np.dtype(t)) for n, t in members]
        self.kernel, self.nano_steps, self.radius = kernel, nano_steps, radius
        self.ref_model = ref_model or name.lower()
        self.cell_dtype = np.dtype([(n, t) for n, t in self.members])
        self.default_cell = np.zeros((), dtype=self.cell_dtype)
        if edge:
            for k, v in edge.items():
                self.default_cell[k] = v

    def step_params(self, final):
        """kernel parameter block for a chunk of sweeps; `final` = the grid is observed afterwards"""
        if self.kernel == capi.KERNEL_LBM_D3Q19:
            # density / velocity are outputs only: store them on the last sweep before an observation
            return np.array([0 if final else 2], dtype=np.int32)
        return None

    @property
    def grid_class(self):
        from .simulator import B200Grid
        return B200Grid

    @property
    def max_fused_sweeps(self):
        """sweeps the kernel family can take per launch (b200geo_update_box_n): 4 for the Jacobi kernels
        (csrc/jacobi_tb.cu), 2 for LBM (csrc/lbm_tb.cu), else 1"""
        if self.kernel in (capi.KERNEL_JACOBI6, capi.KERNEL_JACOBI7, capi.KERNEL_JACOBI27):
            return 4
        return 2 if self.kernel == capi.KERNEL_LBM_D3Q19 else 1

    @property
    def fuses_sweeps(self):
        return self.max_fused_sweeps > 1

    def halo_members(self, width):
        """(members read from the low-side ghost slices, ... high-side) or None = whole cells.
        LBM, ghost width 1: only the populations that cross the face are read from a neighbour's
        plane (csrc/lbm.cu: T* from z-1; B* from z+1, plus SE which the EAST_NOSLIP rule of
        src/examples/latticeboltzmann/main.cpp takes from z+1); a wider ghost zone recomputes the rim
        and needs the neighbours' cells (all that an update reads of them)."""
        if self.kernel == capi.KERNEL_LBM_D3Q19 and width == 1:
            ix = self.member_index
            return ([ix(n) for n in ("T", "TW", "TE", "TN", "TS")],
                    [ix(n) for n in ("B", "BW", "BE", "BN", "BS", "SE")])
        if self.kernel == capi.KERNEL_LBM_D3Q19:
            # the rim is recomputed (two fused sweeps per round): the ghost cells' 19 populations and their state, member by
            # member and in place (no pack kernel, so the transfer overlaps the interior update); density / velocity of a
            # ghost cell are never read
            cell = list(range(19)) + [self.member_index("state")]
            return (cell, cell)
        return None

    @property
    def invariant_members(self):
        """members (indices) the kernel family never rewrites: every sweep reads them from whichever buffer is
        current, so an upload has to put them into BOTH buffers (SerialSimulator initialises both grids,
        parallelization/serialsimulator.h:54-57). LBM: csrc/lbm.cu leaves `state` and the wall cells' density /
        velocity alone."""
        if self.kernel == capi.KERNEL_LBM_D3Q19:
            return [self.member_index(n) for n in ("density", "velocityX", "velocityY", "velocityZ", "state")]
        return []

    @property
    def member_bytes(self):
        return [t.itemsize for _, t in self.members]

    @property
    def wraps(self):
        return self.topology == "torus"

    def member_index(self, name):
        for i, (n, _) in enumerate(self.members):
            if n == name:
                return i
        raise ValueError("no member %r in %s" % (name, self.name))

    def cell_to_bytes(self, cell):
        c = np.zeros((), dtype=self.cell_dtype)
        if isinstance(cell, np.void) or (isinstance(cell, np.ndarray) and cell.dtype == self.cell_dtype):
            c = cell
        elif isinstance(cell, dict):
            c = self.default_cell.copy()
            for k, v in cell.items():
                c[k] = v
        else:  # scalar for single-member cells, else a sequence in registration order
            vals = cell if isinstance(cell, (tuple, list)) else (cell,)
            for (n, _), v in zip(self.members, vals):
                c[n] = v
        return b"".join(np.asarray(c[n]).tobytes() for n, _ in self.members)


class NBodyModel:
    """Binding of BoxCell<FixedArray<LJParticle<REAL>, capacity> > (oracle/models/nbody.h) to the
    n-body kernels: Cube<3>, Moore<3,1>, NANO_STEPS 1; containers of edge `cell_edge` >= cutoff."""

    def __init__(self, name, real, capacity=32, cell_edge=2.5, cutoff=2.5, dt=0.005):
        self.name, self.dim, self.topology, self.radius, self.nano_steps = name, 3, "cube", 1, 1
        self.real = np.dtype(real)
        self.capacity, self.cell_edge, self.cutoff, self.dt = capacity, cell_edge, cutoff, dt
        self.kernel = capi.KERNEL_NBODY
        self.members = [("counts", np.dtype("i4")), ("particles", self.real)]   # the two device arrays
        self.ref_model = "nbody"
        self.wraps = False
        self.fuses_sweeps = False
        self.max_fused_sweeps = 1

    def with_params(self, **kw):
        args = dict(real=self.real, capacity=self.capacity, cell_edge=self.cell_edge, cutoff=self.cutoff, dt=self.dt)
        args.update(kw)
        return NBodyModel(self.name, **args)

    def step_params(self, final):
        return capi.NBodyParams(self.dt, self.cutoff, self.nano_steps)

    def halo_members(self, width):
        return ([0, 1], [0, 1])

    @property
    def grid_class(self):
        from .boxgrid import BoxGrid
        return BoxGrid


class ContainerModel:
    """Binding of ContainerCell<MeshElement<DIM, TORUS>, capacity> (oracle/models/container.h: the mesh element of
    src/examples/voronoi/main.cpp:9-118, neighbours addressed by ID) to the container kernels: Moore<DIM, 1>,
    NANO_STEPS 1."""

    def __init__(self, name, dim=3, topology="cube", capacity=16, max_neighbors=20):
        self.name, self.dim, self.topology, self.radius, self.nano_steps = name, dim, topology, 1, 1
        self.capacity, self.max_neighbors = capacity, max_neighbors
        self.kernel = capi.KERNEL_CONTAINER
        self.ref_model = "container"
        self.wraps = topology == "torus"
        self.fuses_sweeps = False
        self.max_fused_sweeps = 1

    def with_params(self, **kw):
        args = dict(dim=self.dim, topology=self.topology, capacity=self.capacity, max_neighbors=self.max_neighbors)
        args.update(kw)
        return ContainerModel(self.name, **args)

    def step_params(self, final):
        return None

    @property
    def grid_class(self):
        from .containergrid import ContainerGrid
        return ContainerGrid


_F64 = [("temp", "f8")]
_LBM = [(n, "f4") for n in ["C", "N", "E", "W", "S", "T", "B", "NW", "SW", "NE", "SE", "TW", "BW", "TE", "BE",
                            "TN", "BN", "TS", "BS", "density", "velocityX", "velocityY", "velocityZ"]] + [("state", "i4")]

Jacobi6Cube = CellModel("Jacobi6Cube", 3, "cube", _F64, capi.KERNEL_JACOBI6)
Jacobi6Torus = CellModel("Jacobi6Torus", 3, "torus", _F64, capi.KERNEL_JACOBI6)
Jacobi7Cube = CellModel("Jacobi7Cube", 3, "cube", _F64, capi.KERNEL_JACOBI7)
Jacobi7Torus = CellModel("Jacobi7Torus", 3, "torus", _F64, capi.KERNEL_JACOBI7)
Jacobi27Cube = CellModel("Jacobi27Cube", 3, "cube", _F64, capi.KERNEL_JACOBI27)
Jacobi27Torus = CellModel("Jacobi27Torus", 3, "torus", _F64, capi.KERNEL_JACOBI27)
ConwayCube = CellModel("ConwayCube", 2, "cube", [("alive", "u1")], capi.KERNEL_GOL)
ConwayTorus = CellModel("ConwayTorus", 2, "torus", [("alive", "u1")], capi.KERNEL_GOL)
# edge cell = LBMCellF(): C = 1, density = 1 (oracle/models/lbm.h)
LBMCellF = CellModel("LBMCellF", 3, "cube", _LBM, capi.KERNEL_LBM_D3Q19, edge={"C": 1.0, "density": 1.0},
                     ref_model="lbm")

NBodyF = NBodyModel("NBodyF", "f4")
NBodyD = NBodyModel("NBodyD", "f8")

Container2Cube = ContainerModel("Container2Cube", 2, "cube")
Container2Torus = ContainerModel("Container2Torus", 2, "torus")
Container3Cube = ContainerModel("Container3Cube", 3, "cube")
Container3Torus = ContainerModel("Container3Torus", 3, "torus")

ALL = {m.name: m for m in [NBodyF, NBodyD, Container2Cube, Container2Torus, Container3Cube, Container3Torus, Jacobi6Cube, Jacobi6Torus, Jacobi7Cube, Jacobi7Torus, Jacobi27Cube, Jacobi27Torus,
                           ConwayCube, ConwayTorus, LBMCellF]}
