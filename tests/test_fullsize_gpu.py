"""Parity at BASELINE.json's FULL sizes, where running the oracle over the whole grid would take minutes.

The update is local: after s sweeps a cell depends only on the cells within distance s. So the oracle is
run on WINDOWS of the full-size input — the window plus a halo of s cells on every side that is not a
true domain boundary (there the oracle's own edge handling is the real one) — and the GPU's full-size
result must match it bit for bit inside the window. Windows sit at domain corners, faces and in the
interior, across tile / z-chunk seams of the kernels. Game of Life is compared over the whole grid.

  configs[2]  Jacobi 27-point 1024^3 f64     4 and 5 sweeps (two fused launches; plus a one-sweep remainder)
  (and the 7-point kernel at 1024^3, 4 sweeps in one launch)
  configs[3]  LBM D3Q19 512^3 f32, cavity     3 and 6 sweeps (fused pairs + a one-sweep remainder), whole x-y planes at the bottom, in the middle, at the top
  configs[1]  Game of Life 16384^2 u8         8 sweeps, byte kernel and bit-packed path, whole grid
  configs[4]  n-body, 108^3 containers        2 sweeps, the corner window of 8^3 containers (16.6 M particles)
"""
import gc

import numpy as np
import pytest

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Grid

pytestmark = pytest.mark.gpu


def free_enough(gigabytes):
    import torch
    free, _ = torch.cuda.mem_get_info()
    return free >= gigabytes * 2 ** 30


def jacobi_windows(n, s):
    """(z0, y0, x0, dz, dy, dx): corners, a face, interior pieces across the 60 x 28 tile seams and the
    128-plane z-chunk seam of the temporal-blocked kernel"""
    return [(0, 0, 0, 12, 40, 70), (n - 10, n - 33, n - 66, 10, 33, 66), (120, 20, 50, 16, 40, 80),
            (500, n - 30, 0, 9, 30, 64), (n - 7, 300, 471, 7, 31, 130), (255, 0, n - 100, 6, 29, 100)]


@pytest.mark.parametrize("kind,steps", [(27, 4), (27, 5), (7, 4)])
def test_jacobi_1024_cubed_windows(oracle, kind, steps):
    n = 1024
    if not free_enough(20):
        pytest.skip("needs 20 GB of device memory")
    model = models.ALL["Jacobi%dCube" % kind]
    block = synth.jacobi_grid(n, n, 32, seed=5)          # the input repeats every 32 planes along z
    import torch
    grid = B200Grid(model, (n, n, n))
    grid.setEdge(0.5)
    dev_block = torch.from_numpy(block).cuda()           # one trip over PCIe, 32 device-to-device copies
    for z in range(0, n, 32):
        grid.loadMember("temp", dev_block, origin=(0, 0, z), location=capi.CUDA_DEVICE)
    torch.cuda.synchronize()
    del dev_block
    grid.dev.step(model.kernel, steps)
    for z0, y0, x0, dz, dy, dx in jacobi_windows(n, steps):
        lo = [max(0, z0 - steps), max(0, y0 - steps), max(0, x0 - steps)]
        hi = [min(n, z0 + dz + steps), min(n, y0 + dy + steps), min(n, x0 + dx + steps)]
        zs = np.arange(lo[0], hi[0]) % 32
        sub = np.ascontiguousarray(block[zs][:, lo[1]:hi[1], lo[2]:hi[2]])
        want = oracle.jacobi(kind, False, sub, steps, edge=0.5)
        want = want[z0 - lo[0]:z0 - lo[0] + dz, y0 - lo[1]:y0 - lo[1] + dy, x0 - lo[2]:x0 - lo[2] + dx]
        got = grid.saveMember("temp", origin=(x0, y0, z0), dims=(dx, dy, dz))
        assert np.array_equal(got, want), (z0, y0, x0)
    del grid
    gc.collect()


@pytest.mark.parametrize("kind,depth", [(27, 2), (7, 4)])
def test_jacobi_1024_cubed_repeated_fused_launches(kind, depth):
    """the WHOLE 1024^3 grid, not windows: launches of the temporal-blocked kernel (2 fused sweeps of the 27-point stencil,
    4 of the 7-point one) from the same input, every one compared on the device, value by value, with the same number of
    single-sweep launches (whose results the window tests above pin to the oracle). What the windows cannot see — a rare
    ordering bug somewhere in 5000 CTAs — this does (it is how the LBM kernel's was found)."""
    import torch
    n = 1024
    if not free_enough(40):
        pytest.skip("needs 40 GB of device memory")
    model = models.ALL["Jacobi%dCube" % kind]
    block = torch.from_numpy(synth.jacobi_grid(n, n, 32, seed=5)).cuda()

    def fresh():
        grid = B200Grid(model, (n, n, n))
        grid.setEdge(0.5)
        for z in range(0, n, 32):
            grid.loadMember("temp", block, origin=(0, 0, z), location=capi.CUDA_DEVICE)
        return grid

    def values(grid):
        t = torch.empty((n, n, n), dtype=torch.float64, device="cuda")
        grid.saveMember("temp", out=t, location=capi.CUDA_DEVICE)
        return t.view(torch.int64)

    try:
        capi.set_tuning("jacobi.tb", 1)
        grid = fresh()
        grid.dev.step(model.kernel, depth)
        want = values(grid)
        del grid
        capi.set_tuning("jacobi.tb", depth)
        grid = fresh()
        for launch in range(12):
            before = capi.launch_count()
            grid.dev.step(model.kernel, depth)
            assert capi.launch_count() - before == 1
            got = values(grid)
            assert torch.equal(got, want), launch
            del got
            # back to the input: both buffers held it before the first launch, the launch wrote the other one only
            grid.dev.swap()
    finally:
        capi.set_tuning("jacobi.tb", -1)
    del grid, want, block
    gc.collect()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("steps", [3, 6])
def test_lbm_512_cubed_planes(oracle, steps):
    """3 sweeps = one launch of two fused sweeps + one single sweep; 6 = three fused launches. Whole x-y planes: every
    tile seam of the fused kernel in x and y; planes 254..257 straddle a z-chunk seam (multiples of 64)"""
    n = 512
    if not free_enough(32):
        pytest.skip("needs 32 GB of device memory")
    model = models.LBMCellF
    # populations: one noisy 16-plane block repeated along z; wall states: the real cavity
    noise = synth.lbm_grid(n, n, 16, noise=0.01, z0=16, nz_total=n)
    grid = B200Grid(model, (n, n, n))
    for z in range(0, n, 16):
        states = synth.lbm_states(n, n, 16, z, n)
        for m, (name, t) in enumerate(model.members):
            grid.loadMember(name, states if name == "state" else noise[m].view(t), origin=(0, 0, z))
    grid.dev.step(model.kernel, steps)
    for z0, dz in ((0, 4), (254, 4), (n - 4, 4)):
        lo, hi = max(0, z0 - steps), min(n, z0 + dz + steps)
        zs = np.arange(lo, hi) % 16
        raw = np.ascontiguousarray(noise[:, zs])
        raw[23] = synth.lbm_states(n, n, hi - lo, lo, n).view(np.float32)
        want = oracle.lbm(raw, steps)
        for m, (name, t) in enumerate(model.members):
            got = grid.saveMember(name, origin=(0, 0, z0), dims=(n, n, dz))
            assert np.array_equal(got.view(np.int32), want[m, z0 - lo:z0 - lo + dz].view(np.int32)), (name, z0)
    del grid
    gc.collect()


def test_lbm_512_cubed_repeated_fused_launches():
    """The fused LBM kernel hands planes from its sweep-1 warps to its sweep-2 warps through mbarriers; an arrival that
    came before a wall cell's shared-memory loads had been performed made ONE row next to a wall plane wrong in about one
    launch of twenty at this size — and only at this size (profiles/r4o_r4u). 40 launches from the same input, every
    one compared on the device, value by value, with two single-sweep launches (themselves checked against the oracle
    above)."""
    import torch
    n = 512
    if not free_enough(60):
        pytest.skip("needs 60 GB of device memory")
    model = models.LBMCellF
    noise = synth.lbm_grid(n, n, 16, noise=0.01, z0=16, nz_total=n)

    def fresh():
        grid = B200Grid(model, (n, n, n))
        for z in range(0, n, 16):
            states = synth.lbm_states(n, n, 16, z, n)
            for m, (name, t) in enumerate(model.members):
                grid.loadMember(name, states if name == "state" else noise[m].view(t), origin=(0, 0, z))
        return grid

    def populations(grid):
        out = []
        for m in range(19):
            t = torch.empty((n, n, n), dtype=torch.float32, device="cuda")
            grid.saveMember(model.members[m][0], out=t, location=capi.CUDA_DEVICE)
            out.append(t.view(torch.int32))
        return out

    try:
        capi.set_tuning("lbm.tb", 1)
        grid = fresh()
        grid.dev.step(model.kernel, 2)
        want = populations(grid)
        del grid
        capi.set_tuning("lbm.tb", 2)
        grid = fresh()
        for launch in range(40):
            before = capi.launch_count()
            grid.dev.step(model.kernel, 2)      # current -> scratch buffer; the input stays where it is
            assert capi.launch_count() - before == 1
            got = populations(grid)
            for m in range(19):
                assert torch.equal(got[m], want[m]), (launch, model.members[m][0])
            del got
            grid.dev.swap()                     # the next launch starts from the same input
    finally:
        capi.set_tuning("lbm.tb", -1)
    del grid, want
    gc.collect()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("bits", [0, 4])
def test_gol_16384_squared_whole_grid(oracle, bits):
    n, steps = 16384, 8
    capi.set_tuning("gol.bits", bits)
    try:
        block = synth.gol_grid(n, 1024)
        g0 = np.ascontiguousarray(np.tile(block, (16, 1)))
        grid = B200Grid(models.ConwayCube, (n, n))
        grid.loadMember("alive", g0)
        grid.dev.step(capi.KERNEL_GOL, steps)
        got = grid.saveMember("alive")
        assert np.array_equal(got, oracle.gol(False, g0, steps))
    finally:
        capi.set_tuning("gol.bits", -1)
    del grid
    gc.collect()


def test_nbody_108_cubed_containers_corner_window(oracle):
    """the oracle takes container (0,0,0) at the origin, so the window is the corner of the box"""
    n, steps, w = 108, 2, 8
    model = models.NBodyF
    c, p = synth.nbody_cells(n, n, n, vel=2.0, dtype=np.float32)
    assert c.sum() > 16_000_000
    grid = model.grid_class(model, (n, n, n))
    grid.loadCells(c, p)
    grid.dev.step(model.kernel, steps, params=model.step_params(True))
    grid.dev.check()
    gc_, gp = grid.saveCells(origin=(0, 0, 0), dims=(w, w, w))
    h = w + steps
    wc, wp = oracle.nbody(np.ascontiguousarray(c[:h, :h, :h]), np.ascontiguousarray(p[:h, :h, :h]), steps, dt=model.dt)
    assert np.array_equal(gc_, wc[:w, :w, :w])
    assert np.array_equal(gp.view(np.uint8), np.ascontiguousarray(wp[:w, :w, :w]).view(np.uint8))
    del grid
    gc.collect()
