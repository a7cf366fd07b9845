"""The streamed run of StripedSimulator (striping.py::_run_streamed) on the real engine: pinned host arrays in and out,
three CUDA streams, time-skewed b200geo_update_box_n launches, bit-identical to the oracle. The CPU twin (same schedule on
the numpy engine) is tests/test_streamed_run_cpu.py. The tests read the host buffers right after run() returns, with no
synchronisation of their own: ParallelWriter reads through a GridWindow are stream ordered and the contract
(simulator.py::ParallelWriter) says they are complete when run() returns."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,depth,steps,shape,chunks", [
    (27, 2, 8, (160, 96, 128), 5), (27, 2, 7, (96, 40, 70), 4), (7, 4, 12, (128, 64, 96), 4), (7, 1, 5, (64, 33, 50), 8),
    (6, 3, 10, (90, 32, 64), 6)])
def test_streamed_run_on_the_device(kind, depth, steps, shape, chunks):
    """StripedSimulator(stream_io=True).run() on the real engine: pinned host arrays in and out (in place), three
    streams, time-skewed update_box_n launches — bit-identical to the oracle and to the plain run. CPU twin:
    tests/test_streamed_run_cpu.py."""
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    from libgeodecomp_b200 import capi, models, synth
    from libgeodecomp_b200.striping import StripedSimulator
    from oracle import oracle_py
    nz, ny, nx = shape
    data = synth.jacobi_grid(nx, ny, nz, seed=kind)
    want = oracle_py.jacobi(kind, False, data, steps)
    pinned = torch.empty(shape, dtype=torch.float64, pin_memory=True)
    host = {"temp": pinned.numpy()}
    for stream_io in (True, False, True):
        host["temp"][...] = data
        Init, PullWriter = bench.make_plugins(host, host, 2, 0)
        sim = StripedSimulator(Init((nx, ny, nz), steps), models.ALL["Jacobi%dCube" % kind], stream_io=stream_io,
                               stream_depth=depth, stream_chunks=chunks)
        sim.writers = [PullWriter("", 1 << 30)]
        sim.run()     # no synchronisation here: the host buffers must be complete when run() returns
        assert sim.streamed_runs == (1 if stream_io else 0)
        assert np.array_equal(host["temp"], want), "stream_io=%s" % stream_io
        assert np.array_equal(sim.getGrid().saveMember("temp"), want)


@pytest.mark.gpu
def test_streamed_lbm_and_gol_on_the_device():
    """kernel families that take one sweep per launch (b200geo_update_box): LBM D3Q19 (24 members, macroscopics stored
    on the last level only) and the byte Game of Life kernel (2-D: chunks of rows), streamed vs the oracle."""
    import torch
    from libgeodecomp_b200 import capi, models, synth
    from libgeodecomp_b200.simulator import ParallelWriter, SimpleInitializer
    from libgeodecomp_b200.striping import StripedSimulator
    from oracle import oracle_py
    nx, ny, nz, steps = 40, 24, 64, 6
    raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
    pinned = torch.empty(raw.shape, dtype=torch.float32, pin_memory=True)
    host = pinned.numpy()
    host[...] = raw

    class LBMInit(SimpleInitializer):
        def grid(self, target):
            (ox, oy, oz), (dx, dy, dz) = target.boundingBox()
            for m, (name, t) in enumerate(models.LBMCellF.members):
                target.loadMember(name, host[m, oz:oz + dz].view(t), origin=(ox, oy, oz))

    class LBMPull(ParallelWriter):
        def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank, lastCall):
            (ox, oy, oz), (dx, dy, dz) = validRegion
            if event == 2:
                for m, (name, t) in enumerate(models.LBMCellF.members):
                    grid.saveMember(name, origin=(ox, oy, oz), dims=(dx, dy, dz), out=host[m, oz:oz + dz].view(t))

    sim = StripedSimulator(LBMInit((nx, ny, nz), steps), models.LBMCellF, stream_io=True, stream_chunks=4)
    sim.addWriter(LBMPull("", steps))
    sim.run()     # no synchronisation here: the host buffers must be complete when run() returns
    assert sim.streamed_runs == 1
    assert np.array_equal(host.view(np.uint32), oracle_py.lbm(raw, steps).view(np.uint32))

    gx, gy, gsteps = 300, 256, 9
    gol = synth.gol_grid(gx, gy)
    gpinned = torch.empty(gol.shape, dtype=torch.uint8, pin_memory=True)
    ghost = gpinned.numpy()
    ghost[...] = gol

    class GolInit(SimpleInitializer):
        def grid(self, target):
            (ox, oy), (dx, dy) = target.boundingBox()
            target.loadMember("alive", ghost[oy:oy + dy], origin=(ox, oy))

    class GolPull(ParallelWriter):
        def stepFinishedRegion(self, grid, validRegion, globalDimensions, step, event, rank, lastCall):
            (ox, oy), (dx, dy) = validRegion
            if event == 2:
                grid.saveMember("alive", origin=(ox, oy), dims=(dx, dy), out=ghost[oy:oy + dy])

    sim = StripedSimulator(GolInit((gx, gy), gsteps), models.ALL["ConwayCube"], stream_io=True, stream_chunks=8)
    sim.addWriter(GolPull("", gsteps))
    sim.run()     # no synchronisation here: the host buffers must be complete when run() returns
    assert sim.streamed_runs == 1
    assert np.array_equal(ghost, oracle_py.gol(False, gol, gsteps))
