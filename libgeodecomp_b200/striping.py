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
from .simulator import (B200Grid, GridWindow, ParallelWriter, SteererFeedback, STEERER_ALL_DONE, STEERER_INITIALIZED,
                        STEERER_NEXT_STEP, WRITER_ALL_DONE, WRITER_INITIALIZED, WRITER_STEP_FINISHED)


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
        n = len(grid.model.members)
        # need[side]: the members the update reads from the ghost slices on that side. A model may
        # name a subset (LBM with ghost width 1: the five populations that cross the face): those
        # travel in place, member by member, without any pack kernel.
        need = grid.model.halo_members(width)
        self.need = need if need is not None else (list(range(n)), list(range(n)))
        self.packed = n > 1 and need is None
        self._tensors = {}
        self._staging = None

    def _block(self, member, side, kind, which=0):
        blk = self.grid.dev.halo_block(member, side, kind, self.width, which=which)
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
                # device memory for the real engine; the CPU test engine (tests/cpu_engine.py) names its own
                where = getattr(self.grid.engine, "staging_device", "cuda")
                buf = torch.empty(cells * self.grid.model.cell_dtype.itemsize, dtype=torch.uint8, device=where)
                self._staging[(side, kind)] = (st, buf)

    def post(self, which=0):
        """Start filling `width` ghost slices on both PEER sides of the current (which = 0) or the
        scratch (which = 1) buffer from the neighbours' outermost owned slices of the same buffer.
        Returns the pending work handles; the transfer is ordered after everything enqueued so far
        on the current stream and runs concurrently with whatever is enqueued next."""
        dev = self.grid.dev
        if self.world == 1:
            return []
        if self.packed:
            if which != 0:
                raise capi.LogicError("packed halo buffers are taken from the current grid buffer")
            if self._staging is None:
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
            if self.packed:
                ops.append(self.dist.P2POp(self.dist.irecv, self._staging[(side, 1)][1], peer))
            else:
                for m in self.need[side]:
                    ops.append(self.dist.P2POp(self.dist.irecv, self._block(m, side, 1, which), peer))
        for side, peer in ((0, self.low), (1, self.high)):
            if peer is None:
                continue
            if self.packed:
                st, t = self._staging[(side, 0)]
                dev.save_region(st, t, location=capi.CUDA_DEVICE)
                nbytes += t.numel()
                ops.append(self.dist.P2POp(self.dist.isend, t, peer))
            else:
                # our low-side rim lands in the neighbour's HIGH ghost: it wants need[1], and vice versa
                for m in self.need[1 - side]:
                    t = self._block(m, side, 0, which)
                    nbytes += t.numel()
                    ops.append(self.dist.P2POp(self.dist.isend, t, peer))
        self.bytes_per_exchange = nbytes
        return self.dist.batch_isend_irecv(ops) if ops else []

    def finish(self, works):
        """Wait for post() (the current stream waits, not the host), unpack, declare the ghosts valid.
        Call after the swap when post() addressed the scratch buffer."""
        if self.world == 1:
            return
        for work in works:
            work.wait()
        for side, peer in ((0, self.low), (1, self.high)):
            if peer is None:
                continue
            if self.packed:
                st, t = self._staging[(side, 1)]
                self.grid.dev.load_region(st, t, location=capi.CUDA_DEVICE, both=False)
            self.grid.dev.halo_mark_valid(side, self.width)

    def exchange(self):
        """Fill `width` ghost slices on both PEER sides of the current buffer (blocking the stream)."""
        self.finish(self.post(0))


class _CudaStreams:
    """three CUDA streams (host->device, sweeps, device->host) and the events between them; torch is the plumbing"""

    def __init__(self, device):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.up, self.run, self.down = (torch.cuda.Stream(device=self.device) for _ in range(3))
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream(self.device))   # everything here starts after what the caller enqueued so far
        for st in (self.up, self.run, self.down):
            st.wait_event(start)

    @staticmethod
    def handle(stream):
        return stream.cuda_stream

    def record(self, stream):
        ev = self.torch.cuda.Event()
        ev.record(stream)
        return ev

    @staticmethod
    def wait(stream, event):
        stream.wait_event(event)

    def join(self):
        """the caller's current stream continues after all three, and the HOST waits for them: the ParallelWriter
        contract says the host buffers are complete when run() returns"""
        cur = self.torch.cuda.current_stream(self.device)
        for st in (self.up, self.run, self.down):
            cur.wait_event(self.record(st))
        for st in (self.up, self.run, self.down):
            st.synchronize()


class _NoStreams:
    """engines that execute synchronously (tests/cpu_engine.py)"""
    up = run = down = None

    @staticmethod
    def handle(stream):
        return None

    def record(self, stream):
        return None

    @staticmethod
    def wait(stream, event):
        pass

    def join(self):
        pass


class StripedSimulator:
    """DistributedSimulator over equal slabs along the last axis, one rank per GPU.

    Same observable behaviour as the reference's StripingSimulator / HiParSimulator with ghost zone
    width `ghost_width`: every rank initialises only its own slab (Initializer::grid is called
    with a grid whose boundingBox() is the slab plus its ghost planes), writers see the rank's
    own region, results are independent of the number of ranks.
    """

    def __init__(self, initializer, model, rank=0, world=1, ghost_width=1, device=0, dist=None, engine=None,
                 overlap=True, stream_io=False, stream_depth=None, stream_chunks=16):
        """stream_io: let run() pipeline Initializer -> sweeps -> ParallelWriters chunk by chunk along the last axis
        (see _run_streamed) where that is possible; stream_depth = sweeps per launch there (default: what the
        kernel family fuses), stream_chunks = number of chunks the axis is cut into.
        On more than one rank run() can only be streamed when the ghost zones are as wide as the run is long
        (ghost_width >= sweeps of the run): every rank then asks its Initializer for its slab PLUS the ghost zones,
        recomputes the shrinking ghost zone redundantly and needs no exchange at all while the wavefront passes —
        HiParSimulator's ghost zone of width k taken to k = run length (parallelization/hiparsimulator.h:60-66)."""
        if model.kernel == capi.KERNEL_CONTAINER:
            raise capi.LogicError("ContainerCell grids run on one device (B200Simulator): there is no slab partition for them yet")
        self.initializer, self.model, self.rank, self.world = initializer, model, rank, world
        self.overlap = overlap
        self.stream_io, self.stream_depth, self.stream_chunks = stream_io, stream_depth, stream_chunks
        self.device = device
        self.streamed_runs = 0
        self.NANO_STEPS = model.nano_steps
        gdims = tuple(initializer.gridDimensions())
        last = model.dim - 1
        bounds = slab_bounds(gdims[last], world)
        z0, z1 = bounds[rank], bounds[rank + 1]
        # every rank must take the same schedule (the exchanges pair up): decisions that depend on the slab thickness
        # use the thinnest slab of the partition
        self._thinnest = min(bounds[r + 1] - bounds[r] for r in range(world))
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
        self.grid = model.grid_class(model, dims, device=device, ghost_z=ghost_z, z_modes=z_modes, origin=origin,
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
        """nano_steps sweeps with one halo exchange per ghost_width sweeps.

        Rounds of up to ghost_width sweeps follow StripingSimulator::nanoStep's schedule
        (parallelization/stripingsimulator.h:269-286): update the rims, start shipping them, update
        the interior while they travel, wait. Configurations that cannot overlap (packed multi-member
        halos, slabs thinner than two ghost zones) exchange first and step afterwards."""
        done = 0
        w = self.ghost_width
        while done < nano_steps:
            left = nano_steps - done
            if self.world == 1:
                self.grid.dev.step(self.model.kernel, n_steps=left, params=self.model.step_params(True))
                return
            fused = self.overlap and self._can_overlap()
            if self._valid == 0 or (fused and self._valid < w):
                # the first exchange; leftover validity (0 < _valid < w) is topped up the same way, so that every
                # round below is a fused, overlapped one
                self.halo.exchange()
                self._valid = w
            if fused:
                # a round of fewer than w sweeps (the tail of this call) still ships w planes: the ghost zones are
                # w deep again afterwards and the next call starts with an overlapped round, not with an exchange
                n = min(left, w)
                self._overlapped_round(done + n == nano_steps, n)
                done += n
                continue
            n = min(left, self._valid)
            self.grid.dev.step(self.model.kernel, n_steps=n, params=self.model.step_params(done + n == nano_steps))
            self._valid -= n
            done += n

    def _can_overlap(self):
        w, last = self.ghost_width, self.model.dim - 1
        if self.halo.packed or self._thinnest < 2 * w or not hasattr(self.grid.dev, "update_box"):
            return False
        return w == 1 or w <= self.model.max_fused_sweeps

    def _overlapped_round(self, final, sweeps):
        """`sweeps` <= ghost_width sweeps fused per launch; the ghost zones are ghost_width deep again afterwards"""
        g, dev, w, last = self.grid, self.grid.dev, self.ghost_width, self.model.dim - 1
        n = g.dims[last]
        params = self.model.step_params(final)

        def update(a, b):
            if b <= a:
                return
            origin, dim = [0, 0, 0], list(g.dims) + [1] * (3 - len(g.dims))
            origin[last], dim[last] = a, b - a
            dev.update_box(self.model.kernel, origin, dim, params=params, n_sweeps=sweeps)

        if self.model.wraps:
            dev.refresh_ghosts()
        lo = w if self.halo.low is not None else 0
        hi = n - w if self.halo.high is not None else n
        update(0, lo)
        update(hi, n)
        works = self.halo.post(which=1)     # ships the new rims out of the scratch buffer
        update(lo, hi)                      # overlaps with the transfer
        dev.swap()
        self.halo.finish(works)

    def step(self):
        feedback = SteererFeedback()
        self._handleInput(STEERER_NEXT_STEP, feedback)
        self.advance(self.NANO_STEPS)
        self.stepNum += 1
        self._afterStep()

    # ---- streamed run: host -> device -> host as a pipeline along the last axis -------------------------------
    def _stream_plan(self):
        """(sweeps per level, ...) or None when run() cannot be streamed: a Cube topology (periodic images would tie
        the first planes to the last), no Steerers, only ParallelWriters that fire at the very end, and on more than
        one rank ghost zones as wide as the run is long."""
        if not self.stream_io or self.model.wraps or self.steerers:
            return None
        if self.world > 1:
            total = (self.initializer.maxSteps() - self.initializer.startStep()) * self.NANO_STEPS
            if self.ghost_width < total:
                return None
        if not hasattr(self.grid.dev, "update_box") or not hasattr(self.model, "member_index"):
            return None
        steps = self.initializer.maxSteps() - self.initializer.startStep()
        sweeps = steps * self.NANO_STEPS
        if sweeps < 1:
            return None
        first, end = self.initializer.startStep(), self.initializer.maxSteps()
        for w in self.writers:
            # no WRITER_STEP_FINISHED may fall due inside the run: no step s in (first, end) with s % period == 0
            if not isinstance(w, ParallelWriter) or first // w.getPeriod() != (end - 1) // w.getPeriod():
                return None
        depth = self.stream_depth
        if depth is None:
            depth = {capi.KERNEL_JACOBI27: 2, capi.KERNEL_JACOBI6: 4, capi.KERNEL_JACOBI7: 4,
                     capi.KERNEL_LBM_D3Q19: 2}.get(self.model.kernel, 1)
        if not self.model.fuses_sweeps:
            depth = 1
        depth = max(1, min(int(depth), self.model.max_fused_sweeps, sweeps))
        # the shorter remainder level goes FIRST: a level may not be deeper than the one after it (see below)
        levels = ([sweeps % depth] if sweeps % depth else []) + [depth] * (sweeps // depth)
        # every rank must come to the same decision (a rank that falls back to the plain schedule exchanges halos,
        # a streaming one does not): the test is made for the extents of ALL ranks
        mine = None
        for r in range(self.world):
            lo, hi = self._stream_extent(r)
            n = hi - lo
            chunk = max(2 * depth, -(-n // max(1, int(self.stream_chunks))))
            chunk = -(-chunk // depth) * depth     # boxes that are cut off at plane 0 stay whole multiples of the depth
            if n < 2 * chunk:
                return None
            if r == self.rank:
                mine = chunk
        chunk = mine
        return levels, chunk

    def _run_streamed(self, levels, chunk):
        """Time-skewed sweep over chunks of `chunk` planes (rows in 2-D) of the last axis, so that the upload of one
        chunk, the sweeps over the chunks before it and the download of finished planes run at the same time on three
        streams — instead of upload everything, sweep, download everything. The reference's parallel simulators hand
        Initializers and ParallelWriters sub-boxes in the same way (io/initializer.h:38-71, io/parallelwriter.h:92-99);
        results are bit-identical to the plain run (same kernels, same order of arithmetic per cell).

        Level l = levels[l] fused sweeps (one b200geo_update_box_n launch) taking planes from time t_l = sum(levels[:l])
        to t_l + levels[l]; off[l] = sum(levels[:l + 1]). In round c, level l updates planes
        [c * chunk - off[l], (c + 1) * chunk - off[l]): its input reaches up to (c + 1) * chunk - off[l - 1], which
        level l - 1 of the same round has just produced (level 0: up to (c + 1) * chunk, the chunk just uploaded).
        Plane z at time t lives in buffer (number of levels up to t) % 2. Level l + 1 overwrites the buffer level l
        read from, up to plane (c + 1) * chunk - off[l + 1]; round c + 1 will read that buffer from plane
        (c + 1) * chunk - off[l] - levels[l] on — no overlap as long as levels[l + 1] >= levels[l]. All sweeps are on
        one stream in this order, uploads touch only planes no level has reached, downloads only planes that are
        final."""
        g, dev, model = self.grid, self.grid.dev, self.model
        last = model.dim - 1
        own = g.dims[last]
        # the streamed axis: the slab plus, towards a neighbouring slab, the whole ghost zone (plane index s = z - zb)
        zb, ze = self._stream_extent()
        n = ze - zb
        low_peer, high_peer = zb < 0, ze > own
        L = len(levels)
        off = [sum(levels[:l + 1]) for l in range(L)]
        first_nano = [(sum(levels[:l])) % self.NANO_STEPS for l in range(L)]
        streams = _NoStreams() if getattr(g.engine, "synchronous", False) else _CudaStreams(self.device)
        start, last_step = self.initializer.startStep(), self.initializer.maxSteps()
        parity = [0]
        if low_peer:
            dev.halo_mark_valid(0, self.ghost_width)   # filled by the Initializer below, never exchanged
        if high_peer:
            dev.halo_mark_valid(1, self.ghost_width)

        def want(p):
            # which buffer the C ABI calls "current" is a host-side flag; device work already enqueued keeps its pointers
            if parity[0] != p:
                dev.swap()
                parity[0] = p

        def box(a, b):
            origin, dim = [0, 0, 0], list(g.dims) + [1] * (3 - len(g.dims))
            origin[last], dim[last] = a + zb, b - a
            return origin, dim

        def region(a, b):
            o, d = list(g.origin), list(g.dims)
            o[last], d[last] = g.origin[last] + a + zb, b - a
            return (tuple(o), tuple(d))

        uploads = -(-n // chunk)
        rounds = -(-(n + off[-1]) // chunk)
        for c in range(rounds):
            if c < uploads:
                a, b = c * chunk, min((c + 1) * chunk, n)
                want(0)
                window = GridWindow(g, a + zb, b + zb, stream=streams.handle(streams.up))
                self.initializer.grid(window)
                # writers see the rank's own cells only
                wa, wb = max(a + zb, 0), min(b + zb, own)
                if wb > wa:
                    wwin = GridWindow(g, wa, wb, stream=streams.handle(streams.up))
                    for w in self.writers:
                        w.stepFinishedRegion(wwin, region(wa - zb, wb - zb), g.global_dims, start, WRITER_INITIALIZED, self.rank,
                                             wb == own)
                streams.wait(streams.run, streams.record(streams.up))
            for l in range(L):
                a, b = max(c * chunk - off[l], 0), min((c + 1) * chunk - off[l], n)
                # towards a neighbour the ghost zone shrinks by one plane per sweep: nobody delivers fresh planes
                if low_peer:
                    a = max(a, off[l])
                if high_peer:
                    b = min(b, n - off[l])
                if b <= a:
                    continue
                want(l % 2)
                origin, dim = box(a, b)
                dev.update_box(model.kernel, origin, dim, nano_step=first_nano[l], params=model.step_params(l == L - 1),
                               n_sweeps=levels[l], stream=streams.handle(streams.run))
            a, b = max(c * chunk - off[-1], 0), min((c + 1) * chunk - off[-1], n)
            a, b = max(a + zb, 0), min(b + zb, own)      # own planes that have reached the last time step
            if b > a and self.writers:
                streams.wait(streams.down, streams.record(streams.run))
                want(L % 2)
                window = GridWindow(g, a, b, stream=streams.handle(streams.down))
                for w in self.writers:
                    w.stepFinishedRegion(window, region(a - zb, b - zb), g.global_dims, last_step, WRITER_ALL_DONE, self.rank,
                                         b == own)
        want(L % 2)   # the final state is the current buffer from here on
        streams.join()
        for side in (0, 1):
            if (low_peer, high_peer)[side]:
                dev.halo_mark_valid(side, 0)           # the ghost zones are used up
        self.stepNum = last_step
        self.streamed_runs += 1

    def _stream_extent(self, rank=None):
        """planes [lo, hi) of the last axis (local coordinates of that rank) a streamed run sweeps: the slab and its
        PEER ghost zones"""
        rank = self.rank if rank is None else rank
        last = self.model.dim - 1
        bounds = slab_bounds(self.grid.global_dims[last], self.world)
        n = bounds[rank + 1] - bounds[rank]
        if self.world == 1:
            return 0, n
        lo = -self.ghost_width if rank > 0 else 0
        hi = n + self.ghost_width if rank < self.world - 1 else n
        return lo, hi

    def run(self):
        plan = self._stream_plan()
        if plan is not None:
            self._valid = 0
            self._run_streamed(*plan)
            return
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
                if isinstance(w, ParallelWriter):
                    # ParallelWriter::stepFinished(grid, validRegion, globalDims, step, event, rank, lastCall): the rank's slab
                    w.stepFinishedRegion(self.grid, self.grid.boundingBox(), self.grid.global_dims, self.stepNum, event,
                                         self.rank, True)
                else:
                    w.stepFinished(self.grid, self.stepNum, event)

    def _handleInput(self, event, feedback):
        for s in self.steerers:
            if event != STEERER_NEXT_STEP or self.stepNum % s.getPeriod() == 0:
                s.nextStep(self.grid, self.grid.boundingBox(), self.grid.global_dims, self.stepNum, event,
                           self.rank, True, feedback)
