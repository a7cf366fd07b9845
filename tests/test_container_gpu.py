"""Parity of the ContainerCell path (ID-keyed mesh elements: link resolution + gather sweeps of csrc/container.cu,
through the C ABI and B200Simulator) against the oracle and against the fixtures the reference's SerialSimulator
produced: temperatures bit-exact (tolerance 0: the same adds in the same order, one IEEE division), container
contents and order exact."""
import os

import numpy as np
import pytest

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.containergrid import ContainerGrid
from libgeodecomp_b200.simulator import B200Simulator

from test_container_cpu import CellInit, FIELDS, golden_case, golden_cases, model_for, same_bits, two_containers

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[0, 1], ids=["windows", "tiles"])
def container_kernel(request):
    """every test runs with both layouts of the link table ("container.kernel"): 0 = windows of 1024 cargo items sorted
    by neighbour count, 32-bit links into the global value array; 1 = tiles of containers, the neighbourhood's values
    staged in shared memory, 16-bit links"""
    capi.set_tuning("container.kernel", request.param)
    yield request.param
    capi.set_tuning("container.kernel", -1)


def run(box, edge, nd, torus, steps, model=None):
    sim = B200Simulator(CellInit(box, edge, steps), model or model_for(nd, torus))
    sim.run()
    assert sim.getStep() == steps
    return sim


@pytest.mark.parametrize("key", golden_cases()[0])
def test_container_golden_from_the_reference(key):
    box, edge, nd, torus, steps, want = golden_case(golden_cases()[1], key)
    got = run(box, edge, nd, torus, steps).getGrid().saveCells()
    assert same_bits(got["values"], want)
    for n in FIELDS:
        if n != "values":
            assert np.array_equal(got[n], box[n]), n


@pytest.mark.parametrize("dims,nd,torus,edge,steps", [((13, 9, 5), 3, False, True, 11), ((13, 9, 5), 3, True, False, 11),
                                                      ((40, 33, 1), 2, False, True, 20), ((40, 33, 1), 2, True, False, 20),
                                                      ((1, 1, 1), 3, False, False, 3), ((1, 1, 1), 3, True, False, 3),
                                                      ((2, 3, 1), 3, True, False, 4), ((64, 1, 1), 2, False, False, 9),
                                                      ((24, 24, 24), 3, False, False, 6)])
def test_container_bit_exact(oracle, container_kernel, dims, nd, torus, edge, steps):
    box, eb = synth.container_cells(*dims, n_dims=nd, torus=torus, edge=edge, seed=sum(dims))
    sim = run(box, eb, nd, torus, steps)
    got = sim.getGrid().saveCells(fields=("counts", "values"))
    assert same_bits(got["values"], oracle.container(box, steps, n_dims=nd, torus=torus, edge=eb))
    stats = sim.gatherStatistics()[0]
    live = np.arange(box["ids"].shape[-1]) < box["counts"][..., None]
    assert stats["cargo"] == int(box["counts"].sum()) and stats["links"] == int(box["nb_counts"][live].sum())
    # run() re-initialised the grid once and the constructor loaded it once: the links were resolved once, for run()
    assert stats["resolutions"] == 1 and stats["sweeps"] == steps
    assert stats.get("kernel", container_kernel) == container_kernel


@pytest.mark.parametrize("cap,maxnb", [(1, 1), (5, 3), (32, 8), (40, 6), (100, 20), (7, 64), (1500, 2)])
def test_container_capacities(oracle, cap, maxnb):
    box, eb = synth.container_cells(9, 8, 3, cap=cap, maxnb=maxnb, edge=True, seed=cap)
    model = models.Container3Cube.with_params(capacity=cap, max_neighbors=maxnb)
    got = run(box, eb, 3, False, 5, model).getGrid().saveCells(fields=("values",))
    assert same_bits(got["values"], oracle.container(box, 5, edge=eb))


def test_container_duplicate_ids_first_hit_wins(oracle):
    for (dims, nd, torus, edge) in [((9, 8, 4), 3, False, True), ((10, 9, 1), 2, True, False)]:
        box, eb = synth.container_cells(*dims, n_dims=nd, torus=torus, edge=edge, seed=5)
        synth.container_duplicate_ids(box)
        if eb is not None:
            synth.container_duplicate_ids(eb)
        got = run(box, eb, nd, torus, 4).getGrid().saveCells(fields=("values",))
        assert same_bits(got["values"], oracle.container(box, 4, n_dims=nd, torus=torus, edge=eb))
    box = two_containers([(5, 0.0, 0.0, [7]), (7, 10.0, 0.0, [7])], [(7, 20.0, 0.0, [7]), (9, 0.0, 0.0, [7, 5])])
    got = run(box, None, 2, False, 1).getGrid().saveCells(fields=("values",))["values"]
    assert got[0, 0, 0] == 10.0 and got[0, 1, 1] == 10.0


def test_container_errors_match_the_reference():
    """id not found -> std::logic_error when the grid is first updated (neighborhoodadapter.h:63-64), not when it is
    set; unsorted ids and counts beyond the capacity are rejected"""
    box = two_containers([(5, 1.0, 0.0, [4711])], [(6, 1.0, 0.0, [5])])
    sim = B200Simulator(CellInit(box, None, 2), models.Container2Cube)
    with pytest.raises(capi.LogicError, match="4711"):
        sim.run()
    before = sim.getGrid().saveCells(fields=("values",))["values"]
    assert before[0, 0, 0] == 1.0 and before[0, 1, 0] == 1.0     # the grid is unchanged
    dev = capi.DeviceContainerGrid((2, 1, 1), 16, 20, n_dims=2)
    bad = two_containers([(9, 1.0, 0.0, [9]), (5, 1.0, 0.0, [9])], [])
    dev.load(bad)
    with pytest.raises(ValueError):
        dev.step(capi.KERNEL_CONTAINER, 1)
    bad = two_containers([(5, 1.0, 0.0, [5])], [])
    bad["counts"][0, 1] = 17
    dev.load(bad)
    with pytest.raises(IndexError):
        dev.step(capi.KERNEL_CONTAINER, 1)
    with pytest.raises(ValueError):
        dev.load(two_containers([], []), origin=(1, 0, 0), dim=(2, 1, 1))
    with pytest.raises(capi.LogicError):
        dev.step(capi.KERNEL_JACOBI7, 1)
    with pytest.raises(ValueError):
        capi.DeviceContainerGrid((2, 1, 2), 16, 20, n_dims=2)


def test_container_no_neighbours_is_nan(oracle):
    box = two_containers([(5, 1.0, 0.5, [])], [(6, 1.0, 0.0, [5])])
    got = run(box, None, 2, False, 1).getGrid().saveCells(fields=("values",))["values"]
    assert np.isnan(got[0, 0, 0]) and got[0, 1, 0] == 1.0


def test_container_set_get_and_steering_between_steps(oracle):
    """GridBase::set between steps (what a Steerer does): the links are resolved again and the next steps see the new
    element; get returns containers with ascending ids"""
    box, _ = synth.container_cells(6, 5, 1, n_dims=2, seed=12)
    grid = ContainerGrid(models.Container2Cube, (6, 5))
    grid.loadCells(box)
    grid.dev.step(capi.KERNEL_CONTAINER, 3)
    mid = grid.saveCells()
    assert same_bits(mid["values"], oracle.container(box, 3, n_dims=2))
    cell = grid.get((2, 2))
    assert [c["id"] for c in cell] == sorted(c["id"] for c in cell)
    cell = [c for c in cell][:3] + [{"id": 100000, "temperature": 5.0, "influx": 1.0, "neighborIDs": [100000]}]
    grid.set((2, 2), cell)
    after_set = grid.saveCells()
    with pytest.raises(capi.LogicError):       # elements that named the removed ones cannot find them any more
        grid.dev.step(capi.KERNEL_CONTAINER, 1)
    # point every dangling reference at the new element, then carry on
    live_ids = set(int(v) for v in after_set["ids"][after_set["counts"][..., None] > np.arange(16)])
    nb = after_set["nb_ids"]
    dangling = ~np.isin(nb, list(live_ids)) & (np.arange(20) < after_set["nb_counts"][..., None])
    y, x = np.nonzero(dangling.any(axis=(2, 3)))
    assert len(y) > 0 and (abs(y - 2) <= 1).all() and (abs(x - 2) <= 1).all()
    nb[dangling] = 100000
    grid.loadCells(after_set)
    grid.dev.step(capi.KERNEL_CONTAINER, 4)
    assert same_bits(grid.saveCells(fields=("values",))["values"], oracle.container(after_set, 4, n_dims=2))
    assert grid.dev.stats()["resolutions"] == 2


def test_container_boxes_and_device_buffers(oracle):
    """save / load of sub-boxes, and with the arrays in device memory (MemoryLocation::CUDA_DEVICE)"""
    import torch
    box, _ = synth.container_cells(10, 7, 4, seed=8)
    grid = ContainerGrid(models.Container3Cube, (10, 7, 4))
    # load the grid in two halves, the second from device memory
    lo = {n: np.ascontiguousarray(box[n][:, :, :6]) for n in FIELDS}
    hi = {n: torch.from_numpy(np.ascontiguousarray(box[n][:, :, 6:])).cuda() for n in FIELDS}
    grid.loadCells(lo)
    grid.dev.load(hi, origin=(6, 0, 0), dim=(4, 7, 4), location=capi.CUDA_DEVICE)
    grid.dev.step(capi.KERNEL_CONTAINER, 5)
    want = oracle.container(box, 5)
    part = grid.saveCells(origin=(3, 2, 1), dims=(5, 4, 2))
    assert same_bits(part["values"], want[1:3, 2:6, 3:8]) and np.array_equal(part["ids"], box["ids"][1:3, 2:6, 3:8])
    out = torch.zeros((4, 7, 10, 16), dtype=torch.float64, device="cuda")
    grid.dev.save({"values": out}, location=capi.CUDA_DEVICE)
    assert same_bits(out.cpu().numpy(), want)


def test_container_full_size_against_the_oracle(oracle):
    """1.5 M elements, 16 M links (the oracle takes a few seconds): every temperature"""
    box, _ = synth.container_cells(512, 256, 1, n_dims=2, torus=True, seed=2)
    got = run(box, None, 2, True, 4).getGrid().saveCells(fields=("values",))
    assert same_bits(got["values"], oracle.container(box, 4, n_dims=2, torus=True))
