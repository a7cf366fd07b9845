"""Parity of the CUDA path (through the C ABI / B200Simulator) against the oracle.

Bit-exact for every model: Jacobi has no a*b+c site, GoL is integer, and the LBM kernels are
compiled -fmad=false against a -ffp-contract=off oracle (tolerance 0)."""
import numpy as np
import pytest

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Grid, B200Simulator, SimpleInitializer

pytestmark = pytest.mark.gpu


class MemberInit(SimpleInitializer):
    def __init__(self, dims, steps, members, edge=None):
        SimpleInitializer.__init__(self, dims, steps)
        self.members, self.edge = members, edge

    def grid(self, target):
        if self.edge is not None:
            target.setEdge(self.edge)
        for name, arr in self.members.items():
            target.loadMember(name, arr)


def run_jacobi(kind, torus, shape, steps, edge=0.0):
    nz, ny, nx = shape
    data = synth.jacobi_grid(nx, ny, nz)
    model = models.ALL["Jacobi%d%s" % (kind, "Torus" if torus else "Cube")]
    sim = B200Simulator(MemberInit((nx, ny, nz), steps, {"temp": data}, edge=edge), model)
    sim.run()
    assert sim.getStep() == steps
    return data, sim.getGrid().saveMember("temp")


@pytest.mark.parametrize("kind", [6, 7, 27])
@pytest.mark.parametrize("torus", [False, True])
@pytest.mark.parametrize("shape", [(18, 20, 24), (7, 5, 3), (16, 33, 129), (1, 1, 1), (40, 24, 130)])
def test_jacobi_bit_exact(oracle, kind, torus, shape):
    data, got = run_jacobi(kind, torus, shape, 6, edge=0.25)
    want = oracle.jacobi(kind, torus, data, 6, edge=0.25)
    assert np.array_equal(got, want)


@pytest.fixture
def tuning():
    """set b200geo tuning keys for one test, restore the defaults afterwards"""
    keys = []

    def set_(key, value):
        keys.append(key)
        capi.set_tuning(key, value)
    yield set_
    for k in keys:
        capi.set_tuning(k, -1)


@pytest.mark.parametrize("kind", [6, 7, 27])
@pytest.mark.parametrize("torus", [False, True])
@pytest.mark.parametrize("depth", [1, 2, 3, 4])
@pytest.mark.parametrize("rows", [32, 33, 64, 40])
def test_jacobi_temporal_blocking_bit_exact(oracle, tuning, kind, torus, depth, rows):
    """T sweeps per launch (TMA-staged temporal-blocked kernel) == T single sweeps == the oracle;
    7 steps are not a multiple of any depth, so the remainder launches are covered as well."""
    tuning("jacobi.tb", depth)
    tuning("jacobi.tb_rows", rows)
    for shape in [(18, 20, 24), (7, 5, 3), (1, 1, 1), (40, 70, 130), (9, 129, 61), (5, 30, 259)]:
        data, got = run_jacobi(kind, torus, shape, 7, edge=0.25)
        want = oracle.jacobi(kind, torus, data, 7, edge=0.25)
        assert np.array_equal(got, want), shape


def test_jacobi_temporal_blocking_z_chunks(oracle, tuning):
    """several z chunks per column: every chunk warms its pipeline up on its own 2T planes"""
    tuning("jacobi.tb", 4)
    tuning("jacobi.tb_zchunk", 16)
    for kind in (7, 27):
        data, got = run_jacobi(kind, False, (70, 30, 66), 8, edge=-1.5)
        assert np.array_equal(got, oracle.jacobi(kind, False, data, 8, edge=-1.5))


@pytest.mark.parametrize("kind", [6, 7])
@pytest.mark.parametrize("shape,steps", [((128, 128, 128), 7), ((64, 128, 100), 4), ((128, 144, 128), 3), ((64, 128, 128), 2),
                                         ((128, 128, 128), 1)])
def test_jacobi_sm_resident_kernel(oracle, tuning, kind, shape, steps):
    """small Cube grids: all sweeps of a step() call in one cooperative launch, bricks resident in the SMs
    (csrc/jacobi_resident.cu; 8 or 4 planes per brick), odd and even sweep counts; == the streaming kernels == the oracle"""
    tuning("jacobi.tb", 1)          # the streaming path beside it: one launch per sweep
    tuning("jacobi.resident", 1)
    before = capi.launch_count()
    data, got = run_jacobi(kind, False, shape, steps, edge=0.25)
    launches = capi.launch_count() - before
    want = oracle.jacobi(kind, False, data, steps, edge=0.25)
    assert np.array_equal(got, want)
    tuning("jacobi.resident", 0)
    before = capi.launch_count()
    data, plain = run_jacobi(kind, False, shape, steps, edge=0.25)
    assert np.array_equal(plain, want)
    if steps >= 2:
        assert capi.launch_count() - before - launches == steps - 1   # one launch for all sweeps instead of one per sweep


@pytest.mark.parametrize("kind", [7, 27])
def test_jacobi_config1_128cubed_100_steps(oracle, kind):
    """BASELINE.json config 1: 128^3 double, 100 steps, Cube and the example's Torus."""
    data, got = run_jacobi(kind, False, (128, 128, 128), 100)
    assert np.array_equal(got, oracle.jacobi(kind, False, data, 100))


def test_jacobi6_example_torus_128(oracle):
    data, got = run_jacobi(6, True, (128, 128, 128), 30)
    assert np.array_equal(got, oracle.jacobi(6, True, data, 30))


@pytest.mark.parametrize("bits", [0, 4])
@pytest.mark.parametrize("torus", [False, True])
@pytest.mark.parametrize("shape", [(130, 150), (1, 1), (3, 17), (64, 2048), (100, 16), (37, 4099), (33, 4096), (200, 96), (5, 31)])
def test_gol_bit_exact(oracle, tuning, bits, torus, shape):
    """bits = 0: one byte-grid sweep per launch; 4: the bit-packed path (pack, 20 packed sweeps, unpack)
    wherever it applies (a Torus narrower or wider than whole 32-cell words stays on the byte kernel)"""
    tuning("gol.bits", bits)
    ny, nx = shape
    g = synth.gol_grid(nx, ny)
    model = models.ConwayTorus if torus else models.ConwayCube
    sim = B200Simulator(MemberInit((nx, ny), 20, {"alive": g}), model)
    sim.run()
    assert np.array_equal(sim.getGrid().saveMember("alive"), oracle.gol(torus, g, 20))


def test_gol_2048_64_steps(oracle):
    g = synth.gol_grid(2048, 2048)
    sim = B200Simulator(MemberInit((2048, 2048), 64, {"alive": g}), models.ConwayCube)
    sim.run()
    assert np.array_equal(sim.getGrid().saveMember("alive"), oracle.gol(False, g, 64))


@pytest.mark.parametrize("bits", [0, 4])
@pytest.mark.parametrize("shape", [(50, 70), (40, 64), (9, 33)])
def test_gol_alive_edge(oracle, tuning, bits, shape):
    tuning("gol.bits", bits)
    ny, nx = shape
    g = synth.gol_grid(nx, ny)
    sim = B200Simulator(MemberInit((nx, ny), 9, {"alive": g}, edge=1), models.ConwayCube)
    sim.run()
    assert np.array_equal(sim.getGrid().saveMember("alive"), oracle.gol(False, g, 9, edge_alive=1))


def test_gol_packed_and_byte_sweeps_interleave(oracle):
    """a packed multi-sweep call leaves the byte grid (and its ghost ring) exactly as byte sweeps would:
    3 byte sweeps, 10 packed, 2 byte, 5 packed == 20 sweeps"""
    g = synth.gol_grid(300, 77)
    for model, torus in ((models.ConwayCube, False), (models.ConwayTorus, True)):
        nx = 288 if torus else 300
        gg = np.ascontiguousarray(g[:, :nx])
        grid = B200Grid(model, (nx, 77))
        grid.loadMember("alive", gg)
        for n in (3, 10, 2, 5):
            grid.dev.step(capi.KERNEL_GOL, n)
        assert np.array_equal(grid.saveMember("alive"), oracle.gol(torus, gg, 20))


def lbm_members(raw):
    out = {}
    for m, (name, t) in enumerate(models.LBMCellF.members):
        out[name] = raw[m].view(t)
    return out


@pytest.mark.parametrize("variant,block", [(1, 128), (1, 256), (2, 128), (2, 256)])
@pytest.mark.parametrize("shape,steps", [((16, 18, 20), 12), ((5, 4, 3), 5), ((64, 64, 64), 100), ((24, 21, 130), 7), ((9, 7, 300), 4)])
def test_lbm_bit_exact(oracle, tuning, variant, block, shape, steps):
    """variant 1: one row per thread (the default); 2: two rows per thread (odd ny covers the ragged last CTA)"""
    tuning("lbm.variant", variant)
    tuning("lbm.block", block)
    nz, ny, nx = shape
    raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
    sim = B200Simulator(MemberInit((nx, ny, nz), steps, lbm_members(raw)), models.LBMCellF)
    sim.run()
    want = oracle.lbm(raw, steps)
    grid = sim.getGrid()
    for m, (name, t) in enumerate(models.LBMCellF.members):
        got = grid.saveMember(name)
        assert np.array_equal(got.view(np.int32), want[m].view(np.int32)), name


def test_lbm_macroscopics_every_step_equals_lazy(oracle):
    raw = synth.lbm_grid(20, 12, 10, noise=0.01)
    grid = B200Grid(models.LBMCellF, (20, 12, 10))
    for name, arr in lbm_members(raw).items():
        grid.loadMember(name, arr)
    every = np.array([1], dtype=np.int32)
    for _ in range(5):
        grid.dev.step(capi.KERNEL_LBM_D3Q19, 1, params=every)
    want = oracle.lbm(raw, 5)
    for m, (name, t) in enumerate(models.LBMCellF.members):
        assert np.array_equal(grid.saveMember(name).view(np.int32), want[m].view(np.int32)), name
