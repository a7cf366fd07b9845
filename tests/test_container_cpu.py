"""ContainerCell path (ID-keyed mesh elements) without a GPU: the C restatement (oracle/oracle.c:
oracle_container) against the fixtures the REFERENCE's SerialSimulator produced (tests/golden/container.npz) and,
where oracle/_ref is built, against the reference binary live; the semantics the reference's own
storage/test/unit/containercelltest.h pins (ascending ids, insert replaces, lookup order, "id not found"); and the
host side of the Python mirror (ContainerGrid, B200Simulator) on the stand-in engine of tests/cpu_engine.py."""
import os
import re
import sys

import numpy as np
import pytest

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Simulator, SimpleInitializer

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cpu_engine  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = capi.ContainerBox.FIELDS


def golden_cases():
    z = np.load(os.path.join(GOLDEN, "container.npz"))
    return sorted(k[:-len("_in_counts")] for k in z.files if k.endswith("_in_counts")), z


def golden_case(z, key):
    m = re.match(r"container_(\d)(cube|torus)_.*_s(\d+)$", key)
    box = {n: z[key + "_in_" + n] for n in FIELDS}
    edge = {n: z[key + "_edge_" + n] for n in FIELDS} if key + "_edge_counts" in z.files else None
    return box, edge, int(m.group(1)), m.group(2) == "torus", int(m.group(3)), z[key + "_out_values"]


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.mark.parametrize("key", golden_cases()[0])
def test_container_golden(oracle, key):
    box, edge, nd, torus, steps, want = golden_case(golden_cases()[1], key)
    assert same_bits(oracle.container(box, steps, n_dims=nd, torus=torus, edge=edge), want)


@pytest.mark.parametrize("dims,nd,torus,edge", [((7, 3, 4), 3, False, True), ((7, 3, 4), 3, True, False), ((11, 6, 1), 2, False, False),
                                                ((1, 1, 1), 3, True, False), ((3, 1, 1), 2, True, False)])
def test_container_live_reference(oracle, dims, nd, torus, edge):
    if not oracle.have_ref("container"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    box, eb = synth.container_cells(*dims, n_dims=nd, torus=torus, edge=edge, seed=31)
    ref, stats = oracle.run_ref_container(box, 6, n_dims=nd, torus=torus, edge=eb)
    assert stats["elements"] == int(box["counts"].sum())
    assert np.array_equal(ref["ids"], box["ids"]) and np.array_equal(ref["nb_ids"], box["nb_ids"])
    assert same_bits(oracle.container(box, 6, n_dims=nd, torus=torus, edge=eb), ref["values"])


def two_containers(left, right, maxnb=20, cap=16):
    """a 2 x 1 grid (2-D): lists of (id, temperature, influx, neighbour ids)"""
    box = {"counts": np.zeros((1, 2), np.int32), "ids": np.zeros((1, 2, cap), np.int32), "values": np.zeros((1, 2, cap)),
           "influx": np.zeros((1, 2, cap)), "nb_counts": np.zeros((1, 2, cap), np.int32), "nb_ids": np.zeros((1, 2, cap, maxnb), np.int32)}
    for x, cargo in enumerate((left, right)):
        box["counts"][0, x] = len(cargo)
        for s, (i, t, f, nb) in enumerate(cargo):
            box["ids"][0, x, s], box["values"][0, x, s], box["influx"][0, x, s] = i, t, f
            box["nb_counts"][0, x, s] = len(nb)
            box["nb_ids"][0, x, s, :len(nb)] = nb
    return box


def test_lookup_prefers_the_own_container_then_coordbox_order(oracle):
    """NeighborhoodAdapter::operator[] (storage/neighborhoodadapter.h:45-65): id 7 lives in both containers; the element
    in the left container finds its own container's 7, the one in the right container its own as well"""
    box = two_containers([(5, 0.0, 0.0, [7]), (7, 10.0, 0.0, [7])], [(7, 20.0, 0.0, [7]), (9, 0.0, 0.0, [7, 5])])
    out = oracle.container(box, 1, n_dims=2)
    assert out[0, 0, 0] == 10.0 and out[0, 1, 1] == (20.0 + 0.0) / 2
    if oracle.have_ref("container"):
        assert same_bits(oracle.run_ref_container(box, 1, n_dims=2)[0]["values"], out)


def test_id_not_found_is_a_logic_error(oracle):
    """neighborhoodadapter.h:63-64 / containercelltest.h:27-32: an id in none of the 3^DIM containers throws"""
    box = two_containers([(5, 1.0, 0.0, [4711])], [(6, 1.0, 0.0, [5])])
    with pytest.raises(KeyError) as e:
        oracle.container(box, 1, n_dims=2)
    assert e.value.args[0] == 4711
    if oracle.have_ref("container"):
        with pytest.raises(KeyError):
            oracle.run_ref_container(box, 1, n_dims=2)
    # an element two containers away is out of reach as well
    far = synth.container_cells(5, 1, 1, n_dims=2, seed=3)[0]
    far["counts"][0, :] = 1
    far["ids"][0, :, 0] = [10, 20, 30, 40, 50]
    far["nb_counts"][0, :, 0] = 1
    far["nb_ids"][0, :, 0, 0] = [20, 30, 10, 30, 40]
    with pytest.raises(KeyError) as e:
        oracle.container(far, 1, n_dims=2)
    assert e.value.args[0] == 10


def test_no_neighbours_gives_nan_like_the_reference(oracle):
    """temperature / neighborIDs.size() with an empty list is 0 / 0 (src/examples/voronoi/main.cpp:53)"""
    box = two_containers([(5, 1.0, 0.5, [])], [(6, 1.0, 0.0, [5])])
    out = oracle.container(box, 1, n_dims=2)
    assert np.isnan(out[0, 0, 0]) and out[0, 1, 0] == 1.0
    if oracle.have_ref("container"):
        ref = oracle.run_ref_container(box, 1, n_dims=2)[0]["values"]
        assert np.isnan(ref[0, 0, 0]) and ref[0, 1, 0] == 1.0


# ---- host side of the mirror on the stand-in engine --------------------------------------------------------------

class CellInit(SimpleInitializer):
    def __init__(self, box, edge, steps):
        SimpleInitializer.__init__(self, box["counts"].shape[::-1], steps)
        self.box, self.edge = box, edge

    def grid(self, target):
        if self.edge is not None:
            target.setEdge(target._unpack(self.edge))
        target.loadCells(self.box)


def model_for(nd, torus):
    return models.ALL["Container%d%s" % (nd, "Torus" if torus else "Cube")]


@pytest.mark.parametrize("key", golden_cases()[0])
def test_simulator_on_the_stand_in_engine(key):
    box, edge, nd, torus, steps, want = golden_case(golden_cases()[1], key)
    sim = B200Simulator(CellInit(box, edge, steps), model_for(nd, torus), engine=cpu_engine)
    sim.run()
    assert sim.getStep() == steps
    got = sim.getGrid().saveCells(fields=("counts", "values"))
    assert np.array_equal(got["counts"], box["counts"]) and same_bits(got["values"], want)


def test_container_set_get_follow_containercell_insert():
    """containercelltest.h testInsertAndSearch: ids ascend whatever the insertion order, a second insert of an id
    replaces the cargo, a full container throws std::logic_error"""
    from libgeodecomp_b200.containergrid import ContainerGrid
    model = models.Container2Cube.with_params(capacity=5, max_neighbors=4)
    grid = ContainerGrid(model, (3, 2), engine=cpu_engine)
    cell = [{"id": i, "temperature": float(i), "influx": 0.0, "neighborIDs": [1, 2]} for i in (2, 1, 6, 5, 4)]
    cell.append({"id": 4, "temperature": 44.0, "influx": 0.25, "neighborIDs": [6]})
    grid.set((1, 1), cell)
    back = grid.get((1, 1))
    assert [c["id"] for c in back] == [1, 2, 4, 5, 6]
    assert back[2] == {"id": 4, "temperature": 44.0, "influx": 0.25, "neighborIDs": [6]}
    assert grid.get((0, 0)) == []
    with pytest.raises(capi.LogicError):
        grid.set((0, 0), cell + [{"id": 47}])
    with pytest.raises(IndexError):
        grid.set((0, 0), [{"id": 1, "neighborIDs": [1, 2, 3, 4, 5]}])
    grid.setEdge(back[:2])
    assert grid.getEdge() == back[:2]
    with pytest.raises(ValueError):
        ContainerGrid(models.Container3Cube, (3, 2), engine=cpu_engine)


def test_striped_simulator_refuses_container_models():
    from libgeodecomp_b200.striping import StripedSimulator
    box, _ = synth.container_cells(4, 3, 1, n_dims=2)
    with pytest.raises(capi.LogicError):
        StripedSimulator(CellInit(box, None, 1), models.Container2Cube, engine=cpu_engine)


@pytest.mark.parametrize("seed", range(24))
def test_container_random_meshes_against_the_reference(oracle, seed):
    """seeded random grids — dimension, topology, extents down to one container, fill, shortest neighbour lists, an edge
    container, the same id in several containers — C restatement against the reference binary, bit for bit"""
    if not oracle.have_ref("container"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(1000 + seed)
    nd = int(rng.integers(2, 4))
    dims = [int(rng.integers(1, 7)) for _ in range(nd)] + [1] * (3 - nd)
    torus = bool(rng.integers(0, 2))
    edge = bool(rng.integers(0, 2)) and not torus
    box, eb = synth.container_cells(*dims, n_dims=nd, torus=torus, edge=edge, seed=seed, fill=float(rng.uniform(0.2, 1.0)),
                                    min_neighbors=int(rng.integers(0, 3)))
    if rng.integers(0, 3) == 0:
        synth.container_duplicate_ids(box)
        if eb is not None:
            synth.container_duplicate_ids(eb)
    steps = int(rng.integers(1, 9))
    ref, _ = oracle.run_ref_container(box, steps, n_dims=nd, torus=torus, edge=eb)
    got = oracle.container(box, steps, n_dims=nd, torus=torus, edge=eb)
    nan = np.isnan(got)
    assert np.array_equal(nan, np.isnan(ref["values"]))           # elements without neighbours: 0 / 0 on both sides
    assert np.array_equal(np.where(nan, 0, got).view(np.uint64), np.where(nan, 0, ref["values"]).view(np.uint64))
