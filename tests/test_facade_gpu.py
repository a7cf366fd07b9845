"""Runs the C++ drop-in test (tests/facade/facade_test.cpp): the reference's SerialSimulator and
B200Simulator side by side on the same user models, the reference's MockWriter/MockSteerer event
logs, and SoAGrid vs B200Grid region/member byte streams. The binary is built in the container
that has /root/reference (tests/facade/Makefile, also run by __graft_entry__.build()) and travels."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "facade", "_bin", "facade_test")


@pytest.mark.gpu
def test_cpp_facade_drop_in():
    if not os.access(BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/facade_test not built (needs /root/reference at build time)")
    res = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout


GENERIC_BIN = os.path.join(HERE, "facade", "_bin", "generic_test")


@pytest.mark.gpu
def test_generic_device_path_for_unbound_cells():
    """tests/facade/generic_test.cu: user cells without a kernel binding (the reference's TestCell suite of
    cudasimulatortest.h, a Coord<2>-addressed Game of Life, a partially assigned cell, an AoS updateLineX
    model), their own update() compiled by nvcc, bit-identical to SerialSimulator."""
    if not os.access(GENERIC_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/generic_test not built (needs /root/reference at build time)")
    res = subprocess.run([GENERIC_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout


STRIPING_BIN = os.path.join(HERE, "facade", "_bin", "striping_test")


@pytest.mark.gpu
def test_cpp_striping_simulator_drop_in():
    """tests/facade/striping_test.cpp: B200StripingSimulator (one process, slabs round-robin over the GPUs
    present, direct device-to-device halos) bit-identical to SerialSimulator for 1-4 slabs."""
    if not os.access(STRIPING_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/striping_test not built (needs /root/reference at build time)")
    res = subprocess.run([STRIPING_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout


STREAMED_BIN = os.path.join(HERE, "facade", "_bin", "streamed_test")


@pytest.mark.gpu
def test_cpp_streamed_run():
    """tests/facade/streamed_test.cpp: B200Simulator::run() with ParallelWriters only pipelines Initializer, sweeps and
    writers chunk by chunk on three streams (b200streamedrun.h): bit-identical to SerialSimulator for Jacobi 6 / 7 / 27,
    LBM and Game of Life, every cell handed to the writers exactly once; plain schedule where a plugin needs the whole
    grid."""
    if not os.access(STREAMED_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/streamed_test not built (needs /root/reference at build time)")
    res = subprocess.run([STREAMED_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout


GENERIC_SOA_BIN = os.path.join(HERE, "facade", "_bin", "generic_soa_test")


@pytest.mark.gpu
def test_generic_device_path_for_unbound_soa_cells():
    """tests/facade/generic_soa_test.cu: Struct-of-Arrays user cells without a kernel binding, their SoA-signature
    updateLineX() compiled by nvcc with LibFlatArray's generated accessors, bit-identical to SerialSimulator on one
    device and on slab groups."""
    if not os.access(GENERIC_SOA_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/generic_soa_test not built (needs /root/reference at build time)")
    res = subprocess.run([GENERIC_SOA_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout


@pytest.mark.gpu
def test_uniform_element_layout_round_trip():
    """b200geo_grid_create_uniform: members of different widths share one element index; member I/O, region I/O,
    the edge ring and the periodic images behave as in the default layout."""
    from libgeodecomp_b200 import capi
    dim, widths = (37, 9, 5), [8, 4, 4, 1, 2]
    wrap = [[capi.GHOST_WRAP] * 2, [capi.GHOST_EDGE] * 2, [capi.GHOST_WRAP] * 2]
    grids = [capi.DeviceGrid(dim, widths, ghost=(1, 1, 1), ghost_mode=wrap, member_stride=s) for s in (None, 0, 65536)]
    assert grids[1].member_stride % 256 == 0 and grids[2].member_stride == 65536
    rng = np.random.default_rng(5)
    dtypes = {8: np.float64, 4: np.uint32, 2: np.uint16, 1: np.uint8}
    edge = bytes(rng.integers(0, 255, sum(widths), dtype=np.uint8))
    for g in grids:
        g.set_edge(edge)
    data = []
    for m, w in enumerate(widths):
        a = rng.integers(0, 200, dim[::-1]).astype(dtypes[w])
        data.append(a)
        for g in grids:
            g.load_member(m, a)
    for g in grids:
        g.refresh_ghosts()
    # the whole padded box (ghost ring included) must agree with the default layout, member by member
    padded = tuple(d + 2 for d in dim)

    def pull(g, m, w):
        out = np.zeros(padded[::-1], dtype=dtypes[w])
        g.save_member(m, out, origin=(-1, -1, -1), dim=padded)
        capi.sync()
        return out

    for m, w in enumerate(widths):
        want = pull(grids[0], m, w)
        for g in grids[1:]:
            assert np.array_equal(pull(g, m, w), want), "member %d differs in the uniform layout" % m
        assert np.array_equal(want[1:-1, 1:-1, 1:-1], data[m])


STEPPER_BIN = os.path.join(HERE, "facade", "_bin", "stepper_test")


@pytest.mark.gpu
def test_b200_stepper_beside_vanilla_stepper():
    """tests/facade/stepper_test.cpp: B200Stepper as rank 1 of 3 of a ragged StripingPartition, ghost zone
    widths 1-4, 2-D and 3-D, bit-identical to the reference's VanillaStepper and to the whole-space SerialSimulator;
    patches handed to PatchAccepters identical. (The TestCell variant of the reference's own stepper test runs inside
    generic_test.)"""
    if not os.access(STEPPER_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/stepper_test not built (needs /root/reference at build time)")
    res = subprocess.run([STEPPER_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout and "DIFFERENT" not in res.stdout


CHECKPOINT_BIN = os.path.join(HERE, "facade", "_bin", "checkpoint_test")


@pytest.mark.gpu
def test_checkpoints_in_the_mpiio_layout_from_the_device():
    """tests/facade/checkpoint_test.cpp: B200CheckpointWriter / B200ParallelCheckpointWriter write the
    reference's MPI-IO file layout (io/mpiio.h:82-130) from the device grid box by box, B200CheckpointInitializer restarts
    from it; files identical to the ones written from the reference's host grid, restarted runs end where the
    uninterrupted run does."""
    if not os.access(CHECKPOINT_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/checkpoint_test not built (needs /root/reference at build time)")
    res = subprocess.run([CHECKPOINT_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout and "DIFFERENT" not in res.stdout


CONTAINER_BIN = os.path.join(HERE, "facade", "_bin", "container_test")


@pytest.mark.gpu
def test_cpp_container_cell_drop_in():
    """tests/facade/container_test.cpp: ContainerCell<MeshElement, 16> grids (ID-keyed cargo, the cell of the reference's
    Voronoi example) through SerialSimulator and B200Simulator side by side, 2-D / 3-D, Cube / Torus, edge container:
    temperatures bit-identical; B200ContainerGrid set / get / steering writes; std::logic_error for an unknown id."""
    if not os.access(CONTAINER_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/container_test not built (needs /root/reference at build time)")
    res = subprocess.run([CONTAINER_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout and "DIFFERENT" not in res.stdout


VORONOI_BIN = os.path.join(HERE, "facade", "_bin", "voronoi_test")


@pytest.mark.gpu
def test_reference_voronoi_example_drop_in():
    """tests/facade/voronoi_test.cpp: the reference's own src/examples/voronoi/main.cpp (SimpleCell in
    ContainerCell<SimpleCell, 1000>, VoronoiInitializer / VoronoiMesher), compiled unchanged, on SerialSimulator and on
    B200Simulator: one B200GEO_BIND_CARGO line and the simulator type are all that differs; every element bit-identical."""
    if not os.access(VORONOI_BIN, os.X_OK):
        pytest.skip("tests/facade/_bin/voronoi_test not built (needs /root/reference at build time)")
    res = subprocess.run([VORONOI_BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all checks passed" in res.stdout and "DIFFERENT" not in res.stdout


def test_facade_header_has_no_oracle_dependency():
    for name in ("b200containergrid.h", "b200simulator.h", "b200generic.h", "b200genericsoa.h", "b200boxgrid.h", "b200stripingsimulator.h", "b200stepper.h", "b200checkpoint.h"):
        text = open(os.path.join(HERE, "..", "include", "libgeodecomp_b200", name)).read()
        assert "oracle" not in text
