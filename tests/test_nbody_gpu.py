"""Parity of the n-body path (BoxCell container grid + re-bin / force kernels, through the C ABI and
B200Simulator) against the oracle: container occupancy exact, positions and velocities bit-exact
(tolerance 0: same expression trees and neighbour order, -fmad=false vs -ffp-contract=off)."""
import os
import re

import numpy as np
import pytest

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Simulator, SimpleInitializer

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class CellInit(SimpleInitializer):
    def __init__(self, counts, parts, steps):
        SimpleInitializer.__init__(self, counts.shape[::-1], steps)
        self.counts, self.parts = counts, parts

    def grid(self, target):
        target.loadCells(self.counts, self.parts)


def run(model, counts, parts, steps):
    sim = B200Simulator(CellInit(counts, parts, steps), model)
    sim.run()
    assert sim.getStep() == steps
    return sim.getGrid().saveCells()


@pytest.fixture
def nbody_kernel():
    """select the kernel variant ("nbody.kernel") for one test: 1 = re-bin + one-pass force kernels,
    3 = fused re-bin / candidate-list kernel (default), 0 = fused kernel with per-container masks, 8 / 12 = the
    default kernel with runs of 8 / 12 containers per CTA ("nbody.run")"""
    def set_(value):
        capi.set_tuning("nbody.kernel", value if value in (1, 3) else (3 if value in (192, 256) else 0))   # 0 / 8 / 12: the mask kernel
        capi.set_tuning("nbody.run", value if value in (8, 12) else -1)
        capi.set_tuning("nbody.threads", value if value in (192, 256) else -1)   # of the candidate-list kernel (default 224 for float)
    yield set_
    capi.set_tuning("nbody.kernel", -1)
    capi.set_tuning("nbody.run", -1)
    capi.set_tuning("nbody.threads", -1)


@pytest.mark.parametrize("kernel", [0, 1, 3, 8, 12, 192, 256])
@pytest.mark.parametrize("real", [np.float32, np.float64])
@pytest.mark.parametrize("dims,steps,vel,dt", [((6, 5, 4), 10, 8.0, 0.01), ((9, 3, 2), 6, 20.0, 0.02), ((1, 1, 1), 5, 1.0, 0.01),
                                               ((17, 4, 3), 8, 10.0, 0.01), ((8, 8, 8), 10, 0.0, 0.005), ((2, 1, 7), 9, 15.0, 0.01)])
def test_nbody_bit_exact(oracle, nbody_kernel, kernel, real, dims, steps, vel, dt):
    nbody_kernel(kernel)
    c, p = synth.nbody_cells(*dims, vel=vel, dtype=real)
    model = (models.NBodyF if real == np.float32 else models.NBodyD).with_params(dt=dt)
    co, po = run(model, c, p, steps)
    wc, wp = oracle.nbody(c, p, steps, dt=dt)
    assert np.array_equal(co, wc)
    assert np.array_equal(po.view(np.uint8), wp.view(np.uint8))


def test_nbody_config5_shape_32cubed_particles_10_steps(oracle):
    """SURVEY 8(d) C5 parity case: 32^3 lattice sites (14^3 containers of edge 2.5), 10 steps"""
    c, p = synth.nbody_cells(14, 14, 14, vel=1.0, dtype=np.float32)
    co, po = run(models.NBodyF, c, p, 10)
    wc, wp = oracle.nbody(c, p, 10)
    assert np.array_equal(co, wc) and np.array_equal(po.view(np.uint8), wp.view(np.uint8))


def golden_keys():
    z = np.load(os.path.join(GOLDEN, "nbody.npz"))
    return sorted(k[:-len("_in_counts")] for k in z.files if k.endswith("_in_counts"))


@pytest.mark.parametrize("key", golden_keys())
def test_nbody_golden_from_the_reference(key):
    """straight against what the reference's SerialSimulator produced (tests/golden/make_golden.py)"""
    z = np.load(os.path.join(GOLDEN, "nbody.npz"))
    m = re.match(r"nbody_(float32|float64)_.*_s(\d+)_dt([0-9.]+)", key)
    model = (models.NBodyF if m.group(1) == "float32" else models.NBodyD).with_params(dt=float(m.group(3)))
    co, po = run(model, z[key + "_in_counts"], z[key + "_in_parts"], int(m.group(2)))
    assert np.array_equal(co, z[key + "_out_counts"])
    assert np.array_equal(po.view(np.uint8), z[key + "_out_parts"].view(np.uint8))


@pytest.mark.parametrize("kernel", [0, 1, 3, 8, 256])
@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_nbody_dense_containers_overflow_the_candidate_lists(oracle, nbody_kernel, kernel, real):
    """~27 (capacity 32) and ~33 (capacity 48) particles per container: more than 88 candidates pass the
    filter of the fused kernel, so its candidate lists are flushed mid-way"""
    nbody_kernel(kernel)
    for cap, spacing in ((32, 0.84), (48, 0.80)):
        c, p = synth.nbody_cells(5, 4, 3, cap=cap, spacing=spacing, jitter=0.05, vel=3.0, dtype=real)
        assert c.max() > (24 if cap == 32 else 32)
        model = (models.NBodyF if real == np.float32 else models.NBodyD).with_params(capacity=cap, dt=0.002)
        co, po = run(model, c, p, 4)
        wc, wp = oracle.nbody(c, p, 4, dt=0.002)
        assert np.array_equal(co, wc) and np.array_equal(po.view(np.uint8), wp.view(np.uint8))


def test_nbody_other_capacity_and_edge(oracle):
    c, p = synth.nbody_cells(5, 4, 3, cap=48, edge=3.0, vel=6.0, dtype=np.float32)
    model = models.NBodyF.with_params(capacity=48, cell_edge=3.0, cutoff=2.5, dt=0.01)
    co, po = run(model, c, p, 7)
    wc, wp = oracle.nbody(c, p, 7, dt=0.01, cutoff=2.5, edge=3.0)
    assert np.array_equal(co, wc) and np.array_equal(po.view(np.uint8), wp.view(np.uint8))


def test_nbody_capacity_exceeded_is_out_of_range():
    c = np.zeros((1, 1, 2), dtype=np.int32)
    p = np.zeros((1, 1, 2, 32, 6), dtype=np.float32)
    c[0, 0, 0] = c[0, 0, 1] = 20
    p[0, 0, 0, :20, 0] = 2.4
    p[0, 0, 1, :20, 0] = 2.6
    p[0, 0, 1, :20, 3] = -100.0
    p[0, 0, :, :20, 1] = np.linspace(0.1, 2.4, 20)
    p[0, 0, :, :20, 2] = 1.0
    with pytest.raises(IndexError):
        run(models.NBodyF.with_params(dt=0.01, cutoff=0.01), c, p, 2)


def test_boxgrid_set_get_one_container():
    grid = models.NBodyD.grid_class(models.NBodyD, (3, 2, 2))
    cell = np.arange(18, dtype=np.float64).reshape(3, 6)
    grid.set((2, 1, 0), cell)
    assert np.array_equal(grid.get((2, 1, 0)), cell)
    assert grid.get((0, 0, 0)).shape == (0, 6)
    with pytest.raises(IndexError):
        grid.set((0, 0, 0), np.zeros((33, 6)))
    with pytest.raises(ValueError):
        models.NBodyD.grid_class(models.NBodyD, (3, 2, 2)).dev.load(np.zeros(1, np.int32), np.zeros(192), (5, 0, 0), (1, 1, 1))


def test_nbody_cutoff_must_fit_the_container():
    grid = models.NBodyF.grid_class(models.NBodyF, (2, 2, 2))
    with pytest.raises(ValueError):
        grid.dev.step(capi.KERNEL_NBODY, 1, params=capi.NBodyParams(0.005, 3.0, 1))
