"""The oracle (oracle/oracle.c) against (a) the committed golden fixtures the reference's own
SerialSimulator produced (tests/golden/make_golden.py), (b) the reference binaries themselves when
oracle/_ref/ is present, and (c) the known-answer arithmetic of the reference's
storage/test/unit/fixedneighborhoodupdatefunctortest.h (testTorus :109-164, testCube :304-361)."""
import os
import re

import numpy as np
import pytest

from libgeodecomp_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cases(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return sorted(k[:-3] for k in z.files if k.endswith("_in")), z


@pytest.mark.parametrize("key", cases("jacobi")[0])
def test_jacobi_golden(oracle, key):
    z = cases("jacobi")[1]
    m = re.match(r"jacobi(\d+)_(cube|torus)_.*_s(\d+)", key)
    kind, topo, steps = int(m.group(1)), m.group(2), int(m.group(3))
    got = oracle.jacobi(kind, topo == "torus", z[key + "_in"], steps, edge=0.5)
    assert np.array_equal(got, z[key + "_out"])


@pytest.mark.parametrize("key", cases("gol")[0])
def test_gol_golden(oracle, key):
    z = cases("gol")[1]
    m = re.match(r"gol_(cube|torus)_.*_s(\d+)", key)
    got = oracle.gol(m.group(1) == "torus", z[key + "_in"], int(m.group(2)))
    assert np.array_equal(got, z[key + "_out"])


@pytest.mark.parametrize("key", cases("lbm")[0])
def test_lbm_golden(oracle, key):
    z = cases("lbm")[1]
    steps = int(key.rsplit("_s", 1)[1])
    got = oracle.lbm(z[key + "_in"], steps)
    assert np.array_equal(got.view(np.int32), z[key + "_out"].view(np.int32))


def coordinate_field(origin, dims):
    ox, oy, oz = origin
    nx, ny, nz = dims
    x = np.arange(ox, ox + nx, dtype=np.float64)[None, None, :]
    y = np.arange(oy, oy + ny, dtype=np.float64)[None, :, None]
    z = np.arange(oz, oz + nz, dtype=np.float64)[:, None, None]
    return x + y * 1000.0 + z * 1000 * 1000.0


def expected_neighbour_sum(origin, dims, torus):
    """sum over the six von Neumann neighbours of x + 1000 y + 10^6 z with Torus normalisation or
    the reference test's Cube edge cell -1001001 (fixedneighborhoodupdatefunctortest.h:364-385)."""
    f = coordinate_field(origin, dims)
    if torus:
        p = np.pad(f, 1, mode="wrap")
    else:
        p = np.pad(f, 1, mode="constant", constant_values=-1001001.0)
    c = slice(1, -1)
    return (p[:-2, c, c] + p[c, :-2, c] + p[c, c, :-2] + p[c, c, 2:] + p[c, 2:, c] + p[2:, c, c])


@pytest.mark.parametrize("torus", [True, False])
def test_known_answer_neighbour_sums(oracle, torus):
    origin, dims = (10, 20, 30), (200, 100, 50)
    f = coordinate_field(origin, dims)
    got = oracle.jacobi(6, torus, f, 1, edge=-1001001.0)
    assert np.array_equal(got, expected_neighbour_sum(origin, dims, torus) * (1.0 / 6.0))


def test_degenerated_torus(oracle):
    """x extent 1 (testDegeneratedTorus): every x neighbour is the cell itself."""
    f = coordinate_field((0, 0, 0), (1, 10, 20))
    got = oracle.jacobi(6, True, f, 1)
    assert np.array_equal(got, expected_neighbour_sum((0, 0, 0), (1, 10, 20), True) * (1.0 / 6.0))


def test_region_roundtrip(oracle):
    rng = np.random.default_rng(5)
    dims, mb = (13, 7, 5), [8, 4, 1]
    cells = 13 * 7 * 5
    grid = rng.integers(0, 255, size=cells * sum(mb), dtype=np.uint8)
    streaks = [(0, 0, 0, 13), (3, 2, 1, 9), (12, 6, 4, 13), (5, 5, 2, 5)]
    buf = oracle.save_region(grid, dims, mb, streaks)
    assert buf.size == (13 + 6 + 1 + 0) * sum(mb)
    # member-major: first member's cells of all streaks come first
    m0 = grid[:cells * 8].view(np.uint64).reshape(5, 7, 13)
    assert np.array_equal(buf[:20 * 8].view(np.uint64), np.concatenate([m0[0, 0, 0:13], m0[1, 2, 3:9], m0[4, 6, 12:13]]))
    other = np.zeros_like(grid)
    oracle.load_region(other, dims, mb, streaks, buf)
    again = oracle.save_region(other, dims, mb, streaks)
    assert np.array_equal(buf, again)


# ---- live pinning against the reference binaries (present in the build container and shipped to the GPU box)

def need_ref(oracle, model):
    if not oracle.have_ref(model):
        pytest.skip("oracle/_ref/lgd_ref_%s not built (needs /root/reference)" % model)


@pytest.mark.parametrize("kind", [6, 7, 27])
@pytest.mark.parametrize("topo", ["cube", "torus"])
def test_jacobi_vs_reference_binary(oracle, kind, topo):
    need_ref(oracle, "jacobi%d%s" % (kind, topo))
    g = synth.jacobi_grid(40, 24, 18, seed=9)
    out, st = oracle.run_ref("jacobi%d%s" % (kind, topo), g, (40, 24, 18), 11, edge=-3.0 if topo == "cube" else None)
    assert st["simulator"] == "SerialSimulator"
    assert np.array_equal(out.view(np.float64).reshape(g.shape), oracle.jacobi(kind, topo == "torus", g, 11, edge=-3.0))


def test_openmp_reference_matches_serial(oracle):
    need_ref(oracle, "jacobi27cube")
    g = synth.jacobi_grid(32, 32, 16, seed=1)
    a, _ = oracle.run_ref("jacobi27cube", g, (32, 32, 16), 5)
    b, st = oracle.run_ref("jacobi27cube", g, (32, 32, 16), 5, omp=True, threads=4)
    assert st["simulator"] == "OpenMPSimulator"
    assert np.array_equal(a, b)


@pytest.mark.parametrize("topo", ["cube", "torus"])
def test_gol_vs_reference_binary(oracle, topo):
    need_ref(oracle, "conway" + topo)
    g = synth.gol_grid(130, 70, seed=21)
    out, _ = oracle.run_ref("conway" + topo, g, (130, 70, 1), 25)
    assert np.array_equal(out.reshape(g.shape), oracle.gol(topo == "torus", g, 25))


def test_lbm_vs_reference_binary(oracle):
    need_ref(oracle, "lbm")
    raw = synth.lbm_grid(24, 16, 12, noise=0.01)
    out, _ = oracle.run_ref("lbm", raw, (24, 16, 12), 20)
    assert np.array_equal(out.view(np.int32).reshape(raw.shape), oracle.lbm(raw, 20).view(np.int32))


def test_known_answer_vs_reference_binary(oracle):
    """the reference itself on the known-answer field, through its whole SerialSimulator stack"""
    need_ref(oracle, "jacobi6torus")
    f = coordinate_field((0, 0, 0), (20, 10, 6))
    out, _ = oracle.run_ref("jacobi6torus", f, (20, 10, 6), 1)
    assert np.array_equal(out.view(np.float64).reshape(f.shape), expected_neighbour_sum((0, 0, 0), (20, 10, 6), True) * (1.0 / 6.0))


# ---------------------------------------------------------------- n-body in BoxCell containers

def nbody_cases():
    z = np.load(os.path.join(GOLDEN, "nbody.npz"))
    return sorted(k[:-len("_in_counts")] for k in z.files if k.endswith("_in_counts")), z


@pytest.mark.parametrize("key", nbody_cases()[0])
def test_nbody_golden(oracle, key):
    """C restatement == the reference's SerialSimulator over BoxCell<FixedArray<LJParticle, 32>>
    (fixture generated by tests/golden/make_golden.py): container occupancy and every bit of
    every position / velocity, with particles crossing container faces and leaving the Cube."""
    z = nbody_cases()[1]
    m = re.match(r"nbody_(float32|float64)_.*_s(\d+)_dt([0-9.]+)", key)
    co, po = oracle.nbody(z[key + "_in_counts"], z[key + "_in_parts"], int(m.group(2)), dt=float(m.group(3)))
    assert np.array_equal(co, z[key + "_out_counts"])
    assert np.array_equal(po.view(np.uint8), z[key + "_out_parts"].view(np.uint8))
    assert not np.array_equal(co, z[key + "_in_counts"])   # the case does re-bin


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_nbody_live_reference(oracle, real):
    if not oracle.have_ref("nbody"):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    c, p = synth.nbody_cells(5, 4, 6, vel=12.0, seed=99, dtype=real)
    (cr, pr), _ = oracle.run_ref_nbody(c, p, 8, dt=0.01)
    co, po = oracle.nbody(c, p, 8, dt=0.01)
    assert np.array_equal(co, cr) and np.array_equal(po.view(np.uint8), pr.view(np.uint8))


def test_nbody_capacity_exceeded_is_out_of_range(oracle):
    """FixedArray::operator<< throws std::out_of_range("capacity exceeded") (storage/fixedarray.h:77-83)"""
    c = np.zeros((1, 1, 2), dtype=np.int32)
    p = np.zeros((1, 1, 2, 32, 6), dtype=np.float32)
    c[0, 0, 0] = c[0, 0, 1] = 20
    p[0, 0, 0, :20, 0] = 2.4          # all in container 0 ...
    p[0, 0, 1, :20, 0] = 2.6
    p[0, 0, 1, :20, 3] = -100.0       # ... and container 1's particles fly into container 0
    p[0, 0, :, :20, 1] = np.linspace(0.1, 2.4, 20)
    p[0, 0, :, :20, 2] = 1.0
    with pytest.raises(IndexError):
        oracle.nbody(c, p, 2, dt=0.01, cutoff=0.01)
    if oracle.have_ref("nbody"):
        with pytest.raises(IndexError):
            oracle.run_ref_nbody(c, p, 2, dt=0.01, cutoff=0.01)


def test_nbody_neighbour_counts_like_boxcelltest(oracle):
    """storage/test/unit/boxcelltest.h counts neighbours within a distance; here: with a huge dt-free
    probe (dt = 0) nothing moves, so occupancy is a fixed point and positions stay bit-identical."""
    c, p = synth.nbody_cells(4, 4, 4, vel=0.0, dtype=np.float64)
    co, po = oracle.nbody(c, p, 3, dt=0.0)
    assert np.array_equal(co, c) and np.array_equal(po[..., :3], p[..., :3])
