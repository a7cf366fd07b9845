"""TEST-ONLY stand-in for the device engine (the role MockPatchAccepter / MockPatchBuffer play in the
reference's stepper tests, storage/mockpatchaccepter.h): the same DeviceGrid interface as
libgeodecomp_b200.capi, backed by numpy arrays and the oracle, so that the host-side slab / ghost-zone /
halo-exchange logic of libgeodecomp_b200.striping can be exercised with gloo on a machine without a GPU.
Never imported by the product. Cube topologies only."""
import numpy as np
import torch

from libgeodecomp_b200 import capi
from oracle import oracle_py


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
        for buf in (self.bufs if both else [self.arr]):
            buf[member][oz + gz:oz + gz + dim[2], oy:oy + dim[1], ox:ox + dim[0]] = a

    def save_member(self, member, dst, origin=(0, 0, 0), dim=None, location=0, stream=None):
        dim = self.dim if dim is None else dim
        ox, oy, oz = origin
        gz = self.ghost[2]
        dst.reshape(dim[2], dim[1], dim[0])[...] = self.arr[member][oz + gz:oz + gz + dim[2], oy:oy + dim[1], ox:ox + dim[0]]

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
            sl = slice(gz - lo, gz + nz + hi)
            dense = np.ascontiguousarray(self.arr[0][sl])
            kind = {capi.KERNEL_JACOBI6: 6, capi.KERNEL_JACOBI7: 7, capi.KERNEL_JACOBI27: 27}[kernel]
            out = oracle_py.jacobi(kind, False, dense, 1, edge=float(self.edge[0]))
            # the outermost plane on a PEER side had no valid neighbour: drop it
            a, b = (1 if lo else 0), (out.shape[0] - (1 if hi else 0))
            self.bufs[self.cur ^ 1][0][gz - lo + a:gz - lo + b] = out[a:b]
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
        dense = np.ascontiguousarray(self.arr[0][gz + a:gz + b])
        kind = {capi.KERNEL_JACOBI6: 6, capi.KERNEL_JACOBI7: 7, capi.KERNEL_JACOBI27: 27}[kernel]
        out = oracle_py.jacobi(kind, False, dense, n_sweeps, edge=float(self.edge[0]))
        self.bufs[self.cur ^ 1][0][gz + z0:gz + z1] = out[z0 - a:z1 - a]

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


def sync(stream=None):
    pass
