"""Slab partition of the simulation space across one-process-per-GPU ranks, with ghost-zone
(halo) exchange between slab neighbours.

Mirrors, for the regular-grid path:
  StripingPartition                geometry/partitions/stripingpartition.h:57-62 (slabs along the last axis)
  StripingSimulator::nanoStep      parallelization/stripingsimulator.h:269-286, 392-424
                                   (recv outer ghost, send inner ghost, update)
  HiParSimulator / VanillaStepper  parallelization/nesting/vanillastepper.h:93-225: ghost zone width k,
                                   one exchange of k planes every k steps, rim recomputed redundantly
  PatchLink::Accepter/Provider     communication/patchlink.h:127-151, 218-244 (MPI_Isend/Irecv of
                                   saveRegion buffers) -> here torch.distributed P2P ops (NCCL over
                                   NVLink) straight on the device-resident planes: with slabs along the
                                   last axis a member's k ghost planes are ONE contiguous block, so no
                                   pack/unpack kernel runs at all.
"""
import numpy as np

from . import capi
from .simulator import (B200Grid, SteererFeedback, STEERER_ALL_DONE, STEERER_INITIALIZED, STEERER_NEXT_STEP,
                        WRITER_ALL_DONE, WRITER_INITIALIZED, WRITER_STEP_FINISHED)


def slab_bounds(extent, ranks):
    """Equal-weight StripingPartition: rank r owns planes [b[r], b[r+1])."""
    return [(extent * r) // ranks for r in range(ranks + 1)]


class HaloExchanger:
    """Ghost-slice exchange for one slab. `dist` is torch.distributed (or None for world 1).

    Single-member grids (Jacobi, GoL): a member's `width` boundary slices are one contiguous block of
    device memory, sent and received in place — no pack kernel, no staging copy.
    Multi-member grids (LBM): the boundary region is serialised with ONE saveRegion launch into a
    member-major device buffer — byte for byte the payload PatchLink::Accepter::put ships
    (communication/patchlink.h:127-151) — sent as one message per neighbour, and loadRegion'ed into
    the ghost slices on the other side (Provider::get, patchlink.h:218-244)."""

    def __init__(self, grid, rank, world, width, periodic, dist=None):
        self.grid, self.rank, self.world, self.width, self.periodic = grid, rank, world, width, periodic
        self.dist = dist
        self.low = rank - 1 if rank > 0 else (world - 1 if periodic and world > 1 else None)
        self.high = rank + 1 if rank < world - 1 else (0 if periodic and world > 1 else None)
        self.bytes_per_exchange = 0
        self.packed = len(grid.model.members) > 1
        self._tensors = {}
        self._staging = None

    def _block(self, member, side, kind):
        blk = self.grid.dev.halo_block(member, side, kind, self.width)
        key = getattr(blk, "ptr", None)
        if key is None:
            return blk.as_tensor()
        t = self._tensors.get(key)
        if t is None:
            t = self._tensors[key] = blk.as_tensor()
        return t

    def _slice_streaks(self, side, kind):
        """whole padded slices (ghost columns included) as streaks {x, y, z, endX}"""
        g = self.grid
        nx, gx = g.dims[0], g.ghost[0]
        w = self.width
        if g.model.dim == 3:
            ny, gy, nz = g.dims[1], g.ghost[1], g.dims[2]
            z0 = (0 if side == 0 else nz - w) if kind == 0 else (-w if side == 0 else nz)
            zs, ys = np.meshgrid(np.arange(z0, z0 + w), np.arange(-gy, ny + gy), indexing="ij")
        else:
            ny = g.dims[1]
            y0 = (0 if side == 0 else ny - w) if kind == 0 else (-w if side == 0 else ny)
            ys = np.arange(y0, y0 + w)
            zs = np.zeros_like(ys)
        n = ys.size
        st = np.empty((n, 4), dtype=np.int32)
        st[:, 0], st[:, 1], st[:, 2], st[:, 3] = -gx, ys.reshape(-1), zs.reshape(-1), nx + gx
        return st

    def _setup_staging(self):
        import torch
        self._staging = {}
        for side in (0, 1):
            for kind in (0, 1):
                st = self._slice_streaks(side, kind)
                cells = int((st[:, 3] - st[:, 0]).sum())
                buf = torch.empty(cells * self.grid.model.cell_dtype.itemsize, dtype=torch.uint8, device="cuda")
                self._staging[(side, kind)] = (st, buf)

    def exchange(self):
        """Fill `width` ghost slices on both PEER sides of the current buffer."""
        dev, w = self.grid.dev, self.width
        if self.world == 1:
            return
        if self.packed and self._staging is None:
            self._setup_staging()
        ops = []
        nbytes = 0
        # post receives first, then sends; one batched group = one NCCL launch. Messages between one
        # pair of ranks match in posting order, and on a periodic 2-rank ring both neighbours are the
        # same rank: a neighbour's LOW-side send (posted first) belongs in our HIGH ghost, so the
        # high-side receive is posted first.
        for side, peer in ((1, self.high), (0, self.low)):
            if peer is None:
                continue
            t = self._staging[(side, 1)][1] if self.packed else self._block(0, side, 1)
            ops.append(self.dist.P2POp(self.dist.irecv, t, peer))
        for side, peer in ((0, self.low), (1, self.high)):
            if peer is None:
                continue
            if self.packed:
                st, t = self._staging[(side, 0)]
                dev.save_region(st, t, location=capi.CUDA_DEVICE)
            else:
                t = self._block(0, side, 0)
            nbytes += t.numel()
            ops.append(self.dist.P2POp(self.dist.isend, t, peer))
        if ops:
            for work in self.dist.batch_isend_irecv(ops):
                work.wait()
        self.bytes_per_exchange = nbytes
        for side, peer in ((0, self.low), (1, self.high)):
            if peer is None:
                continue
            if self.packed:
                st, t = self._staging[(side, 1)]
                dev.load_region(st, t, location=capi.CUDA_DEVICE, both=False)
            dev.halo_mark_valid(side, w)


class StripedSimulator:
    """DistributedSimulator over equal slabs along the last axis, one rank per GPU.

    Same observable behaviour as the reference's StripingSimulator / HiParSimulator with ghost zone
    width `ghost_width`: every rank initialises only its own slab (Initializer::grid is called
    with a grid whose boundingBox() is the slab plus its ghost planes), writers see the rank's
    own region, results are independent of the number of ranks.
    """

    def __init__(self, initializer, model, rank=0, world=1, ghost_width=1, device=0, dist=None, engine=None):
        self.initializer, self.model, self.rank, self.world = initializer, model, rank, world
        self.NANO_STEPS = model.nano_steps
        gdims = tuple(initializer.gridDimensions())
        last = model.dim - 1
        bounds = slab_bounds(gdims[last], world)
        z0, z1 = bounds[rank], bounds[rank + 1]
        if z1 - z0 < ghost_width and world > 1:
            raise ValueError("slab thinner than the ghost zone")
        periodic = model.wraps
        if world == 1:
            z_modes = None
            ghost_z = max(model.radius, 1)
        else:
            edge = capi.GHOST_EDGE
            z_modes = [capi.GHOST_PEER if (rank > 0 or periodic) else edge,
                       capi.GHOST_PEER if (rank < world - 1 or periodic) else edge]
            ghost_z = ghost_width
        dims = list(gdims)
        dims[last] = z1 - z0
        origin = [0] * model.dim
        origin[last] = z0
        self.grid = B200Grid(model, dims, device=device, ghost_z=ghost_z, z_modes=z_modes, origin=origin,
                             global_dims=gdims, engine=engine)
        self.ghost_width = ghost_z if world > 1 else 1
        self.halo = HaloExchanger(self.grid, rank, world, self.ghost_width, periodic, dist=dist)
        self.dist = dist
        initializer.grid(self.grid)
        self.stepNum = initializer.startStep()
        self.writers, self.steerers = [], []
        self._valid = 0

    def addWriter(self, writer):
        self.writers.append(writer)

    def addSteerer(self, steerer):
        self.steerers.append(steerer)

    def getStep(self):
        return self.stepNum

    def getGrid(self):
        return self.grid

    def advance(self, nano_steps):
        """nano_steps sweeps with one halo exchange per ghost_width sweeps."""
        done = 0
        while done < nano_steps:
            if self.world > 1 and self._valid == 0:
                self.halo.exchange()
                self._valid = self.ghost_width
            n = min(nano_steps - done, self._valid) if self.world > 1 else nano_steps - done
            self.grid.dev.step(self.model.kernel, n_steps=n, params=self.model.step_params(done + n == nano_steps))
            if self.world > 1:
                self._valid -= n
            done += n

    def step(self):
        feedback = SteererFeedback()
        self._handleInput(STEERER_NEXT_STEP, feedback)
        self.advance(self.NANO_STEPS)
        self.stepNum += 1
        self._afterStep()

    def run(self):
        self.initializer.grid(self.grid)
        self._valid = 0
        self.stepNum = self.initializer.startStep()
        feedback = SteererFeedback()
        self._handleInput(STEERER_INITIALIZED, feedback)
        self._handleOutput(WRITER_INITIALIZED)
        maxSteps = self.initializer.maxSteps()
        while self.stepNum < maxSteps and not feedback.simulationEnded():
            self._handleInput(STEERER_NEXT_STEP, feedback)
            n = maxSteps - self.stepNum
            for p in [w.getPeriod() for w in self.writers] + [s.getPeriod() for s in self.steerers]:
                n = min(n, p - self.stepNum % p)
            n = max(1, n)
            self.advance(n * self.NANO_STEPS)
            self.stepNum += n
            self._afterStep()
        self._handleInput(STEERER_ALL_DONE, feedback)

    def _afterStep(self):
        event = WRITER_ALL_DONE if self.stepNum == self.initializer.maxSteps() else WRITER_STEP_FINISHED
        self._handleOutput(event)

    def _handleOutput(self, event):
        for w in self.writers:
            if event != WRITER_STEP_FINISHED or self.stepNum % w.getPeriod() == 0:
                # ParallelWriter::stepFinished(grid, validRegion, globalDims, step, event, rank, lastCall)
                w.stepFinished(self.grid, self.stepNum, event)

    def _handleInput(self, event, feedback):
        for s in self.steerers:
            if event != STEERER_NEXT_STEP or self.stepNum % s.getPeriod() == 0:
                s.nextStep(self.grid, self.grid.boundingBox(), self.grid.global_dims, self.stepNum, event,
                           self.rank, True, feedback)
