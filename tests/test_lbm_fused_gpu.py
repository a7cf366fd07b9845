"""The LBM kernel that fuses two sweeps per launch (csrc/lbm_tb.cu) against the oracle: bit-exact (tolerance 0; the
kernels are compiled -fmad=false, the oracle -ffp-contract=off) for every tile shape, z chunking, ragged tiles, odd sweep
counts, wall cells anywhere in the domain (the reference's cell accepts any state in any cell,
src/examples/latticeboltzmann/main.cpp:62-93) and boxes inside the grid (b200geo_update_box_n)."""
import numpy as np
import pytest

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Grid

pytestmark = pytest.mark.gpu

M = models.LBMCellF


@pytest.fixture
def tuning():
    keys = []

    def set_(key, value):
        keys.append(key)
        capi.set_tuning(key, value)
    yield set_
    for k in keys:
        capi.set_tuning(k, -1)


def load(grid, raw):
    for m, (name, t) in enumerate(M.members):
        grid.loadMember(name, raw[m].view(t))


def assert_equal(grid, want, what=""):
    for m, (name, t) in enumerate(M.members):
        got = grid.saveMember(name)
        assert np.array_equal(got.view(np.int32), want[m].view(np.int32)), "%s member %s" % (what, name)


# (lbm.tb_warps, lbm.tb_rows): all warps do both sweeps (two CTA-wide barriers per plane; 14 rows x 3 stages, 14 x 2, 16 x 2,
# 8 x 2 with two CTAs per SM) / sweep 1 and sweep 2 on different warps linked by mbarriers (12 rows x 3 stages = the default, 14 x 2,
# 12 x 2, 11 x 3, 10 x 4, 10 x 3)
VARIANTS = [(0, 14), (0, 142), (0, 16), (0, 8), (1, 123), (1, 142), (1, 122), (1, 113), (1, 104), (1, 103)]


@pytest.mark.parametrize("warps,rows", VARIANTS)
@pytest.mark.parametrize("zchunk", [0, 8])
@pytest.mark.parametrize("shape,steps", [((16, 18, 20), 12), ((5, 4, 3), 5), ((3, 40, 70), 2), ((21, 31, 97), 7),
                                         ((40, 29, 33), 4), ((2, 2, 2), 6), ((64, 64, 64), 30)])
def test_lbm_fused_bit_exact(oracle, tuning, warps, rows, zchunk, shape, steps):
    """the lid-driven cavity of configs[3] at small sizes: grids narrower and wider than a tile, ragged last tiles in x
    and y, z chunks of 8 planes (several CTAs along z, first-sweep planes recomputed at the seams)"""
    tuning("lbm.tb", 2)
    tuning("lbm.tb_warps", warps)
    tuning("lbm.tb_rows", rows)
    tuning("lbm.tb_zchunk", zchunk)
    nz, ny, nx = shape
    raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
    grid = B200Grid(M, (nx, ny, nz))
    load(grid, raw)
    before = capi.launch_count()
    grid.dev.step(capi.KERNEL_LBM_D3Q19, steps)
    assert capi.launch_count() - before == (steps + 1) // 2, "two sweeps per launch, one for an odd tail"
    assert_equal(grid, oracle.lbm(raw, steps))


@pytest.mark.parametrize("warps,rows", VARIANTS)
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_lbm_fused_walls_anywhere(oracle, tuning, warps, rows, seed):
    """every wall state in random cells of the domain — inside tiles, on tile seams, in the ring of first-sweep cells
    that two CTAs compute — on top of the cavity's faces: both sweeps of a launch go through the wall rules"""
    tuning("lbm.tb_warps", warps)
    tuning("lbm.tb_rows", rows)
    tuning("lbm.tb_zchunk", 8)
    nx, ny, nz = 75, 37, 19
    raw = synth.lbm_grid(nx, ny, nz, noise=0.02, seed=seed)
    rng = np.random.RandomState(seed)
    states = raw[23].view(np.int32).copy()
    mask = rng.rand(nz, ny, nx) < 0.3
    states[mask] = rng.randint(0, 7, size=int(mask.sum()))
    raw[23] = states.view(np.float32)
    for tb in (1, 2):
        tuning("lbm.tb", tb)
        grid = B200Grid(M, (nx, ny, nz))
        load(grid, raw)
        grid.dev.step(capi.KERNEL_LBM_D3Q19, 5)
        assert_equal(grid, oracle.lbm(raw, 5), "tb %d" % tb)


@pytest.mark.parametrize("warps", [0, 1])
def test_lbm_fused_equals_one_sweep_per_launch(tuning, warps):
    """no oracle in between: the two schedules of the device path agree on a grid of several hundred tiles"""
    nx, ny, nz = 200, 150, 70
    raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
    results = []
    tuning("lbm.tb_warps", warps)
    for tb in (1, 2):
        tuning("lbm.tb", tb)
        grid = B200Grid(M, (nx, ny, nz))
        load(grid, raw)
        grid.dev.step(capi.KERNEL_LBM_D3Q19, 9)
        results.append([grid.saveMember(name) for name, _ in M.members])
    for (name, _), a, b in zip(M.members, *results):
        assert np.array_equal(a.view(np.int32), b.view(np.int32)), name


def test_lbm_macroscopics_modes_fused(oracle, tuning):
    """density / velocity: stored by the second sweep of the last launch (mode 0), by every launch (mode 1: the grid after
    the call is the same), never (mode 2: what the first load put there stays)"""
    tuning("lbm.tb", 2)
    nx, ny, nz = 40, 20, 12
    raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
    want = oracle.lbm(raw, 6)
    for mode in (0, 1):
        grid = B200Grid(M, (nx, ny, nz))
        load(grid, raw)
        grid.dev.step(capi.KERNEL_LBM_D3Q19, 6, params=np.array([mode], dtype=np.int32))
        assert_equal(grid, want, "mode %d" % mode)
    grid = B200Grid(M, (nx, ny, nz))
    load(grid, raw)
    grid.dev.step(capi.KERNEL_LBM_D3Q19, 6, params=np.array([2], dtype=np.int32))
    for m, (name, t) in enumerate(M.members):
        ref = raw[m] if name in ("density", "velocityX", "velocityY", "velocityZ") else want[m]
        assert np.array_equal(grid.saveMember(name).view(np.int32), ref.view(np.int32)), name


@pytest.mark.parametrize("origin,dim", [((0, 0, 0), (50, 30, 20)), ((3, 5, 2), (40, 17, 9)), ((31, 13, 7), (3, 2, 1)),
                                        ((1, 0, 11), (49, 30, 9)), ((17, 1, 0), (33, 28, 20))])
@pytest.mark.parametrize("warps", [0, 1])
def test_lbm_update_box_two_sweeps(oracle, tuning, origin, dim, warps):
    """b200geo_update_box_n(n_sweeps = 2) on boxes inside the grid: the box holds the cells of time step 2 (the ring of
    first-sweep cells around it is computed from the grid, cells outside the simulation area stay the edge cell),
    everything else in the scratch buffer is untouched"""
    tuning("lbm.tb_zchunk", 8)
    tuning("lbm.tb_warps", warps)
    nx, ny, nz = 50, 30, 20
    raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
    want = oracle.lbm(raw, 2)
    grid = B200Grid(M, (nx, ny, nz))
    load(grid, raw)   # both buffers hold time step 0
    grid.dev.update_box(capi.KERNEL_LBM_D3Q19, origin, dim, n_sweeps=2, params=np.array([0], dtype=np.int32))
    grid.dev.swap()
    (x0, y0, z0), (dx, dy, dz) = origin, dim
    for m, (name, t) in enumerate(M.members):
        got = grid.saveMember(name).view(np.int32)
        ref = raw[m].view(np.int32).copy()
        ref[z0:z0 + dz, y0:y0 + dy, x0:x0 + dx] = want[m].view(np.int32)[z0:z0 + dz, y0:y0 + dy, x0:x0 + dx]
        assert np.array_equal(got, ref), name


def test_lbm_update_box_n_refuses_three_sweeps():
    grid = B200Grid(M, (8, 8, 8))
    with pytest.raises(ValueError):
        grid.dev.update_box(capi.KERNEL_LBM_D3Q19, (0, 0, 0), (8, 8, 8), n_sweeps=3)
