"""TEST-ONLY stand-in for the device engine (the role MockPatchAccepter / MockPatchBuffer play in the
reference's stepper tests, storage/mockpatchaccepter.h): the same DeviceGrid interface as
libgeodecomp_b200.capi, backed by numpy arrays and the oracle, so that the host-side slab / ghost-zone /
halo-exchange logic of libgeodecomp_b200.striping can be exercised with gloo on a machine without a GPU.
Never imported by the product. Jacobi (one f64 member; Cube and Torus), LBM D3Q19 (24 members, Cube) and
slabs of BoxCell containers (n-body; counts + particles, one ghost plane of containers per side)."""
import numpy as np
import torch

from libgeodecomp_b200 import capi
from oracle import oracle_py


synchronous = True        # no streams: striping.StripedSimulator._run_streamed issues everything in order
staging_device = "cpu"   # where striping.HaloExchanger allocates its packed halo buffers for this engine

_JACOBI = {capi.KERNEL_JACOBI6: 6, capi.KERNEL_JACOBI7: 7, capi.KERNEL_JACOBI27: 27}


class _Block:
    def __init__(self, arr):
        self.arr = arr

    def as_tensor(self):
        return torch.from_numpy(self.arr.reshape(-1).view(np.uint8))


class DeviceGrid:
    def __init__(self, dim, member_bytes, ghost=(1, 1, 1), ghost_mode=None, device=0):
        self.dim, self.ghost, self.mode = tuple(dim), tuple(ghost), ghost_mode
        self.member_bytes = list(member_bytes)
        nx, ny, nz = self.dim
        gz = self.ghost[2]
        dt = {8: np.float64, 4: np.float32, 1: np.uint8}
        # z-padded only; x/y edges are handled by the oracle's constant edge; two buffers like the engine
        self.bufs = [[np.zeros((nz + 2 * gz, ny, nx), dtype=dt[b]) for b in member_bytes] for _ in range(2)]
        self.cur = 0
        self.edge = [np.zeros((), dtype=dt[b]) for b in member_bytes]
        self.valid = [0, 0]

    @property
    def arr(self):
        return self.bufs[self.cur]

    def set_edge(self, cell, stream=None):
        off = 0
        for m, b in enumerate(self.member_bytes):
            self.edge[m] = np.frombuffer(cell[off:off + b], dtype=self.arr[m].dtype)[0]
            off += b
            gz = self.ghost[2]
            for buf in self.bufs:
                if gz and self.mode[2][0] == capi.GHOST_EDGE:
                    buf[m][:gz] = self.edge[m]
                if gz and self.mode[2][1] == capi.GHOST_EDGE:
                    buf[m][-gz:] = self.edge[m]

    def load_member(self, member, src, origin=(0, 0, 0), dim=None, location=0, both=True, stream=None):
        dim = self.dim if dim is None else dim
        ox, oy, oz = origin
        gz = self.ghost[2]
        a = np.asarray(src).reshape(dim[2], dim[1], dim[0])
        if a.dtype != self.arr[member].dtype:      # e.g. the int32 state member of the LBM cell: keep the bits
            a = np.ascontiguousarray(a).view(self.arr[member].dtype)
        for buf in (self.bufs if both else [self.arr]):
            buf[member][oz + gz:oz + gz + dim[2], oy:oy + dim[1], ox:ox + dim[0]] = a

    def save_member(self, member, dst, origin=(0, 0, 0), dim=None, location=0, stream=None):
        dim = self.dim if dim is None else dim
        ox, oy, oz = origin
        gz = self.ghost[2]
        out = dst.reshape(dim[2], dim[1], dim[0])
        if out.dtype != self.arr[member].dtype:
            out = out.view(self.arr[member].dtype)
        out[...] = self.arr[member][oz + gz:oz + gz + dim[2], oy:oy + dim[1], ox:ox + dim[0]]

    def _sweeps(self, kernel, planes, n):
        """n sweeps of the kernel family over the padded planes [planes) of the current buffer, all sides treated
        as domain boundaries by the oracle; returns one dense array per member"""
        if kernel in _JACOBI:
            dense = np.ascontiguousarray(self.arr[0][planes])
            # Torus: x and y wrap inside the oracle; its z wrap only reaches the n outermost planes of the range,
            # which are ghost planes of a slab (dropped by the callers) or, on a single rank, the true images
            torus = self.mode[0][0] == capi.GHOST_WRAP
            return [oracle_py.jacobi(_JACOBI[kernel], torus, dense, n, edge=float(self.edge[0]))]
        if kernel == capi.KERNEL_LBM_D3Q19:
            raw = np.stack([np.ascontiguousarray(self.arr[m][planes]).view(np.float32) for m in range(24)])
            out = oracle_py.lbm(raw, n)
            return [out[m].view(self.arr[m].dtype) for m in range(24)]
        raise capi.LogicError("no kernel bound for this id in the CPU mock")

    def _region(self, streaks, buf, save, both=False):
        """member-major (de)serialisation of streaks {x, y, z, endX}; this mock keeps no x / y ghost cells, so the
        parts of a streak outside the interior read as zero and are skipped on load"""
        st = np.asarray(streaks, dtype=np.int64).reshape(-1, 4)
        count = int((st[:, 3] - st[:, 0]).sum())
        raw = buf.numpy() if hasattr(buf, "numpy") else np.asarray(buf)
        raw = raw.reshape(-1).view(np.uint8)
        nx, ny, nz = self.dim
        gz = self.ghost[2]
        moff = 0
        for m, b in enumerate(self.member_bytes):
            block = raw[moff:moff + count * b].view(self.arr[m].dtype)
            pos = 0
            for x0, y, z, x1 in st:
                n = int(x1 - x0)
                a, e = max(int(x0), 0), min(int(x1), nx)
                if 0 <= y < ny and a < e:
                    seg = slice(pos + a - int(x0), pos + e - int(x0))
                    if save:
                        block[seg] = self.arr[m][z + gz, y, a:e]
                    else:
                        for bufs in (self.bufs if both else [self.arr]):
                            bufs[m][z + gz, y, a:e] = block[seg]
                elif save:
                    block[pos:pos + n] = 0
                pos += n
            moff += count * b

    def save_region(self, streaks, buf, location=0, stream=None):
        self._region(streaks, buf, True)

    def load_region(self, streaks, buf, location=0, both=True, stream=None):
        self._region(streaks, buf, False, both)

    def step(self, kernel, n_steps=1, first_nano_step=0, params=None, stream=None):
        gz, nz = self.ghost[2], self.dim[2]
        for _ in range(n_steps):
            lo = hi = 0
            for side in (0, 1):
                if self.mode[2][side] == capi.GHOST_PEER:
                    if self.valid[side] < 1:
                        raise capi.LogicError("ghost zone exhausted")
                    if side == 0:
                        lo = self.valid[0]
                    else:
                        hi = self.valid[1]
            outs = self._sweeps(kernel, slice(gz - lo, gz + nz + hi), 1)
            # the outermost plane on a PEER side had no valid neighbour: drop it
            for m, out in enumerate(outs):
                a, b = (1 if lo else 0), (out.shape[0] - (1 if hi else 0))
                self.bufs[self.cur ^ 1][m][gz - lo + a:gz - lo + b] = out[a:b]
            self.cur ^= 1
            for side in (0, 1):
                if self.mode[2][side] == capi.GHOST_PEER:
                    self.valid[side] -= 1

    def update_box(self, kernel, origin, dim, nano_step=0, params=None, stream=None, n_sweeps=1):
        """n_sweeps of the box [origin, origin + dim) (whole planes): current -> scratch, no swap"""
        gz, nz = self.ghost[2], self.dim[2]
        assert tuple(origin[:2]) == (0, 0) and tuple(dim[:2]) == self.dim[:2]
        z0, z1 = origin[2], origin[2] + dim[2]
        # input: the box plus n_sweeps planes on either side, cut off where the simulation space ends
        a = z0 - n_sweeps if (z0 - n_sweeps >= 0 or self.mode[2][0] == capi.GHOST_PEER) else 0
        b = z1 + n_sweeps if (z1 + n_sweeps <= nz or self.mode[2][1] == capi.GHOST_PEER) else nz
        assert a >= -self.valid[0] and b <= nz + self.valid[1], "ghost zone exhausted"
        for m, out in enumerate(self._sweeps(kernel, slice(gz + a, gz + b), n_sweeps)):
            self.bufs[self.cur ^ 1][m][gz + z0:gz + z1] = out[z0 - a:z1 - a]

    def swap(self):
        self.cur ^= 1

    def refresh_ghosts(self, stream=None):
        pass

    def halo_block(self, member, side, kind, width=1, which=0):
        gz, nz = self.ghost[2], self.dim[2]
        if kind == 0:
            z = gz if side == 0 else gz + nz - width
        else:
            z = gz - width if side == 0 else gz + nz
        return _Block(self.bufs[self.cur ^ which][member][z:z + width])

    def halo_mark_valid(self, side, width):
        self.valid[side] = width

    def stats_enable(self, on=True):
        pass


class DeviceBoxGrid:
    """capi.DeviceBoxGrid on host arrays: containers in the interchange format (counts [z][y][x], particles
    [z][y][x][capacity][6]) with one ghost plane of containers per z side; the sweep is the oracle's, run on the slab
    plus its valid ghost planes at the slab's place in the global container lattice (container origins enter the
    position check of the re-bin, storage/boxcell.h:150-174)"""

    def __init__(self, dim, capacity, real_bytes, cell_edge, cell_origin=(0, 0, 0), ghost_mode=None, device=0):
        self.dim, self.capacity, self.cell_edge = tuple(dim), int(capacity), float(cell_edge)
        self.cell_origin, self.mode = tuple(cell_origin), ghost_mode
        nx, ny, nz = self.dim
        real = {4: np.float32, 8: np.float64}[real_bytes]
        self.bufs = [[np.zeros((nz + 2, ny, nx), dtype=np.int32), np.zeros((nz + 2, ny, nx, capacity, 6), dtype=real)]
                     for _ in range(2)]
        self.cur = 0
        self.valid = [0, 0]
        self.overflow = False

    def load(self, counts, particles, origin=(0, 0, 0), dim=None, location=0, both=True, stream=None):
        dim = self.dim if dim is None else dim
        (ox, oy, oz), (dx, dy, dz) = origin, dim
        for buf in (self.bufs if both else [self.bufs[self.cur]]):
            buf[0][oz + 1:oz + 1 + dz, oy:oy + dy, ox:ox + dx] = np.asarray(counts).reshape(dz, dy, dx)
            buf[1][oz + 1:oz + 1 + dz, oy:oy + dy, ox:ox + dx] = np.asarray(particles).reshape(dz, dy, dx, self.capacity, 6)

    def save(self, counts, particles, origin=(0, 0, 0), dim=None, location=0, stream=None):
        dim = self.dim if dim is None else dim
        (ox, oy, oz), (dx, dy, dz) = origin, dim
        c, p = self.bufs[self.cur]
        counts.reshape(dz, dy, dx)[...] = c[oz + 1:oz + 1 + dz, oy:oy + dy, ox:ox + dx]
        particles.reshape(dz, dy, dx, self.capacity, 6)[...] = p[oz + 1:oz + 1 + dz, oy:oy + dy, ox:ox + dx]

    def step(self, kernel, n_steps=1, first_nano_step=0, params=None, stream=None):
        if kernel != capi.KERNEL_NBODY or not isinstance(params, capi.NBodyParams):
            raise capi.LogicError("a BoxCell grid steps with KERNEL_NBODY and NBodyParams")
        nz = self.dim[2]
        for _ in range(n_steps):
            lo = hi = 0
            for side in (0, 1):
                if self.mode[2][side] == capi.GHOST_PEER:
                    if self.valid[side] < 1:
                        raise capi.LogicError("ghost zone exhausted")
                    lo, hi = (1, hi) if side == 0 else (lo, 1)
            c, p = self.bufs[self.cur]
            planes = slice(1 - lo, 1 + nz + hi)
            org = (self.cell_origin[0], self.cell_origin[1], self.cell_origin[2] - lo)
            try:
                oc, op = oracle_py.nbody(c[planes], p[planes], 1, dt=params.dt, cutoff=params.cutoff, edge=self.cell_edge,
                                         origin=org)
            except IndexError:      # reported by check(), like the device's sticky overflow flag
                self.overflow = True
                return
            nc, npart = self.bufs[self.cur ^ 1]
            nc[1:1 + nz], npart[1:1 + nz] = oc[lo:lo + nz], op[lo:lo + nz]
            self.cur ^= 1
            for side in (0, 1):
                if self.mode[2][side] == capi.GHOST_PEER:
                    self.valid[side] -= 1

    def check(self, stream=None):
        if self.overflow:
            raise IndexError("capacity exceeded")

    def halo_block(self, member, side, kind, width=1, which=0):
        if width != 1:
            raise ValueError("BoxCell grids have a ghost zone of one container")
        nz = self.dim[2]
        z = (1 if side == 0 else nz) if kind == 0 else (0 if side == 0 else nz + 1)
        return _Block(self.bufs[self.cur ^ which][member][z:z + 1])

    def halo_mark_valid(self, side, width):
        self.valid[side] = width

    def stats_enable(self, on=True):
        pass


def sync(stream=None):
    pass


class DeviceContainerGrid:
    """capi.DeviceContainerGrid on host arrays: the slot store in the interchange format, the sweep is the oracle's
    (ContainerCell grids of ID-keyed mesh elements)."""
    FIELDS = capi.ContainerBox.FIELDS
    _TYPES = (np.int32, np.int32, np.float64, np.float64, np.int32, np.int32)

    def __init__(self, dim, capacity, max_neighbors, n_dims=3, ghost_mode=None, device=0):
        if n_dims not in (2, 3) or (n_dims == 2 and dim[2] != 1):
            raise ValueError("n_dims must be 2 or 3")
        self.dim, self.capacity, self.max_neighbors, self.n_dims = tuple(dim), int(capacity), int(max_neighbors), n_dims
        self.torus = bool(ghost_mode is not None and ghost_mode[0][0] == capi.GHOST_WRAP)
        nx, ny, nz = self.dim
        self.store = self._empty((nz, ny, nx))
        self.edge = self._empty((1, 1, 1))
        self.sweeps, self.resolutions, self.dirty = 0, 0, True

    def _empty(self, shape):
        per = [(), (self.capacity,), (self.capacity,), (self.capacity,), (self.capacity,), (self.capacity, self.max_neighbors)]
        return {n: np.zeros(shape + p, dtype=t) for n, t, p in zip(self.FIELDS, self._TYPES, per)}

    @staticmethod
    def _get(arrays, name):
        return arrays.get(name) if isinstance(arrays, dict) else arrays[DeviceContainerGrid.FIELDS.index(name)]

    def load(self, arrays, origin=(0, 0, 0), dim=None, location=0, stream=None):
        dim = self.dim if dim is None else dim
        (ox, oy, oz), (dx, dy, dz) = origin, dim
        for n in self.FIELDS:
            a = self._get(arrays, n)
            if a is None:
                raise ValueError("a load needs every array of the box")
            tgt = self.store[n][oz:oz + dz, oy:oy + dy, ox:ox + dx]
            tgt[...] = np.asarray(a).reshape(tgt.shape)
        self.dirty = True

    def save(self, arrays, origin=(0, 0, 0), dim=None, location=0, stream=None):
        dim = self.dim if dim is None else dim
        (ox, oy, oz), (dx, dy, dz) = origin, dim
        counts = self.store["counts"][oz:oz + dz, oy:oy + dy, ox:ox + dx]
        live = np.arange(self.capacity) < counts[..., None]
        for n in self.FIELDS:
            a = self._get(arrays, n)
            if a is None:
                continue
            src = self.store[n][oz:oz + dz, oy:oy + dy, ox:ox + dx]
            if n != "counts":
                src = np.where(live if src.ndim == 4 else live[..., None], src, 0)
            a.reshape(src.shape)[...] = src

    def set_edge(self, arrays):
        for n in self.FIELDS:
            self.edge[n].reshape(-1)[...] = np.asarray(self._get(arrays, n)).reshape(-1)
        self.dirty = True

    def get_edge(self, arrays):
        for n in self.FIELDS:
            a = self._get(arrays, n)
            if a is not None:
                a.reshape(-1)[...] = self.edge[n].reshape(-1)

    def step(self, kernel, n_steps=1, first_nano_step=0, params=None, stream=None):
        if kernel != capi.KERNEL_CONTAINER:
            raise capi.LogicError("a ContainerCell grid steps with KERNEL_CONTAINER")
        if n_steps == 0:
            return
        c = self.store["counts"]
        if c.max(initial=0) > self.capacity or self.edge["counts"].max() > self.capacity:
            raise IndexError("ContainerCell capacity exeeded")
        ids = self.store["ids"]
        live = np.arange(self.capacity) < c[..., None]
        if ((np.diff(ids, axis=-1) <= 0) & live[..., 1:]).any():
            raise ValueError("the ids of a container must ascend")
        box = {n: (v if self.n_dims == 3 else v[0]) for n, v in self.store.items()}
        edge = {n: v.reshape(v.shape[3:]) if n != "counts" else v.reshape(1) for n, v in self.edge.items()}
        try:
            out = oracle_py.container(box, n_steps, n_dims=self.n_dims, torus=self.torus, edge=edge)
        except KeyError as e:
            raise capi.LogicError("id not found: could not find id %s in neighborhood" % e.args[0])
        self.store["values"][...] = out.reshape(self.store["values"].shape)
        self.resolutions += self.dirty
        self.dirty = False
        self.sweeps += n_steps

    def stats_enable(self, on=True):
        pass

    def stats(self):
        live = np.arange(self.capacity) < self.store["counts"][..., None]
        return {"cargo": int(self.store["counts"].sum()), "links": int(self.store["nb_counts"][live].sum()), "resolutions": self.resolutions,
                "sweeps": self.sweeps}
