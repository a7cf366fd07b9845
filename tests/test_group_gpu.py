"""Slab groups (b200geo_group_*): the slabs of one simulation space stepped by ONE host thread with the
rim-first schedule and direct device-to-device halo copies. Results must be bit-identical to the
single-domain oracle for every slab count, ghost-zone width and topology.

The slabs are spread round-robin over the devices present: on a 1-GPU box all slabs live on device 0
(the copies are then plain device-to-device copies, the schedule is the same), on a multi-GPU box
neighbouring slabs sit on different GPUs and the halos travel over NVLink."""
import numpy as np
import pytest

from libgeodecomp_b200 import capi, models, synth
from libgeodecomp_b200.simulator import B200Grid
from libgeodecomp_b200.striping import slab_bounds

pytestmark = pytest.mark.gpu


def make_slabs(model, gdims, n_slabs, ghost, members, edge=None):
    """B200Grid slabs along the last axis, initialised like StripedSimulator does: every slab is
    handed its own planes AND its ghost planes (Initializer::grid with boundingBox = slab + ghosts)."""
    last = model.dim - 1
    bounds = slab_bounds(gdims[last], n_slabs)
    ndev = max(1, capi.device_count())
    grids = []
    for r in range(n_slabs):
        z0, z1 = bounds[r], bounds[r + 1]
        peer, edge_mode = capi.GHOST_PEER, capi.GHOST_EDGE
        modes = [peer if (r > 0 or model.wraps) else edge_mode, peer if (r < n_slabs - 1 or model.wraps) else edge_mode]
        dims = list(gdims)
        dims[last] = z1 - z0
        origin = [0] * model.dim
        origin[last] = z0
        g = B200Grid(model, dims, device=r % ndev, ghost_z=ghost, z_modes=modes, origin=origin, global_dims=gdims)
        if edge is not None:
            g.setEdge(edge)
        sl = [slice(None)] * model.dim
        sl[0] = slice(z0, z1)        # arrays are [z][y][x] (or [y][x]): the last axis comes first
        for name, arr in members.items():
            g.loadMember(name, np.ascontiguousarray(arr[tuple(sl)]), origin=origin)
        grids.append(g)
    return grids, bounds


def gather(grids, name):
    return np.concatenate([g.saveMember(name) for g in grids], axis=0)


JACOBI_CASES = [
    # kind, topology, slabs, ghost width, steps, (nz, ny, nx)
    (7, "Cube", 2, 1, 6, (24, 18, 40)),
    (27, "Cube", 3, 3, 8, (30, 12, 36)),
    (27, "Torus", 2, 2, 7, (16, 10, 34)),
    (6, "Torus", 3, 1, 5, (12, 9, 20)),
    (7, "Torus", 4, 4, 9, (32, 8, 70)),
    (27, "Cube", 4, 2, 9, (40, 70, 130)),
    (7, "Cube", 1, 1, 5, (9, 8, 7)),
    (27, "Cube", 2, 4, 6, (8, 5, 33)),      # slabs as thin as the ghost zone: no overlap possible, exchange-then-step
]


@pytest.mark.parametrize("kind,topo,slabs,ghost,steps,shape", JACOBI_CASES)
def test_group_jacobi_bit_exact(oracle, kind, topo, slabs, ghost, steps, shape):
    nz, ny, nx = shape
    data = synth.jacobi_grid(nx, ny, nz, seed=kind)
    model = models.ALL["Jacobi%d%s" % (kind, topo)]
    if slabs == 1:
        grids = [B200Grid(model, (nx, ny, nz))]
        grids[0].setEdge(0.75)
        grids[0].loadMember("temp", data)
    else:
        grids, _ = make_slabs(model, (nx, ny, nz), slabs, ghost, {"temp": data}, edge=0.75)
    group = capi.SlabGroup([g.dev for g in grids], periodic=model.wraps and slabs > 1)
    group.step(model.kernel, steps)
    group.sync()
    want = oracle.jacobi(kind, topo == "Torus", data, steps, edge=0.75)
    assert np.array_equal(gather(grids, "temp"), want)
    if slabs > 1:
        st = group.stats()
        assert st["exchanges"] >= steps // ghost and st["bytes"] > 0
    # a second call continues from the ghosts the last round left valid
    group.step(model.kernel, 3)
    group.sync()
    assert np.array_equal(gather(grids, "temp"), oracle.jacobi(kind, topo == "Torus", data, steps + 3, edge=0.75))
    group.close()


@pytest.mark.parametrize("slabs,ghost,steps", [(2, 1, 9), (3, 1, 4), (2, 2, 7)])
def test_group_lbm_bit_exact(oracle, slabs, ghost, steps):
    """ghost width 1: only the populations that cross a face travel (in place, member by member);
    ghost width 2: whole cells, exchange-then-step"""
    nz, ny, nx = 18, 12, 20
    raw = synth.lbm_grid(nx, ny, nz, noise=0.01)
    model = models.LBMCellF
    members = {name: raw[m].view(t) for m, (name, t) in enumerate(model.members)}
    grids, _ = make_slabs(model, (nx, ny, nz), slabs, ghost, members)
    group = capi.SlabGroup([g.dev for g in grids])
    group.step(model.kernel, steps)
    group.sync()
    want = oracle.lbm(raw, steps)
    for m, (name, t) in enumerate(model.members):
        assert np.array_equal(gather(grids, name).view(np.int32), want[m].view(np.int32)), name
    group.close()


@pytest.mark.parametrize("torus,slabs,ghost", [(False, 2, 1), (True, 3, 1), (False, 3, 2)])
def test_group_gol_2d_bit_exact(oracle, torus, slabs, ghost):
    """2-D grids: slabs along y, the halo is whole rows"""
    ny, nx = 90, 128
    g0 = synth.gol_grid(nx, ny)
    model = models.ConwayTorus if torus else models.ConwayCube
    grids, _ = make_slabs(model, (nx, ny), slabs, ghost, {"alive": g0})
    group = capi.SlabGroup([g.dev for g in grids], periodic=torus)
    group.step(model.kernel, 11)
    group.sync()
    assert np.array_equal(gather(grids, "alive"), oracle.gol(torus, g0, 11))
    group.close()


def test_group_invalidate_after_host_write(oracle):
    """a Steerer-style host write into one slab makes the neighbours' ghost copies stale: invalidate()
    forces a full exchange before the next sweep"""
    nz, ny, nx = 20, 9, 16
    data = synth.jacobi_grid(nx, ny, nz)
    model = models.Jacobi7Cube
    grids, bounds = make_slabs(model, (nx, ny, nz), 2, 1, {"temp": data})
    group = capi.SlabGroup([g.dev for g in grids])
    group.step(model.kernel, 2)
    group.sync()
    mid = oracle.jacobi(7, False, data, 2)
    patched = mid.copy()
    patched[bounds[1] - 1] = 3.5            # the plane the high slab reads as its ghost
    z = bounds[1] - 1
    grids[0].loadMember("temp", np.full((1, ny, nx), 3.5), origin=(0, 0, z))
    group.invalidate()
    group.step(model.kernel, 3)
    group.sync()
    assert np.array_equal(gather(grids, "temp"), oracle.jacobi(7, False, patched, 3))
    group.close()


@pytest.mark.parametrize("real", [np.float32, np.float64])
@pytest.mark.parametrize("slabs,dims,steps", [(2, (6, 5, 8), 7), (3, (5, 4, 9), 6), (4, (4, 3, 4), 5), (1, (3, 3, 3), 4)])
def test_boxgroup_nbody_bit_exact(oracle, real, slabs, dims, steps):
    """slabs of BoxCell containers along z: per sweep every slab pulls its neighbours' boundary container planes
    (the particle migration message) and re-bins / updates; occupancy and particles bit-identical to the oracle"""
    nx, ny, nz = dims
    c, p = synth.nbody_cells(nx, ny, nz, vel=12.0, dtype=real)
    model = (models.NBodyF if real == np.float32 else models.NBodyD).with_params(dt=0.01)
    bounds = slab_bounds(nz, slabs)
    ndev = max(1, capi.device_count())
    grids = []
    for r in range(slabs):
        z0, z1 = bounds[r], bounds[r + 1]
        modes = None if slabs == 1 else [capi.GHOST_PEER if r > 0 else capi.GHOST_EDGE,
                                         capi.GHOST_PEER if r < slabs - 1 else capi.GHOST_EDGE]
        g = model.grid_class(model, (nx, ny, z1 - z0), device=r % ndev, z_modes=modes, origin=(0, 0, z0), global_dims=dims)
        g.loadCells(np.ascontiguousarray(c[z0:z1]), np.ascontiguousarray(p[z0:z1]), origin=(0, 0, z0))
        grids.append(g)
    group = capi.BoxSlabGroup([g.dev for g in grids])
    group.step(model.step_params(True), steps)
    group.sync()
    wc, wp = oracle.nbody(c, p, steps, dt=0.01)
    gc = np.concatenate([g.saveCells()[0] for g in grids], axis=0)
    gp = np.concatenate([g.saveCells()[1] for g in grids], axis=0)
    assert np.array_equal(gc, wc)
    assert np.array_equal(gp.view(np.uint8), wp.view(np.uint8))
    if slabs > 1:
        assert group.stats()["exchanges"] == steps
    group.close()


def test_group_rejects_mismatched_slabs():
    a = capi.DeviceGrid((8, 8, 8), [8], ghost=(1, 1, 1),
                        ghost_mode=[[0, 0], [0, 0], [capi.GHOST_EDGE, capi.GHOST_PEER]])
    b = capi.DeviceGrid((8, 6, 8), [8], ghost=(1, 1, 1),
                        ghost_mode=[[0, 0], [0, 0], [capi.GHOST_PEER, capi.GHOST_EDGE]])
    with pytest.raises(ValueError):
        capi.SlabGroup([a, b])
    c = capi.DeviceGrid((8, 8, 8), [8], ghost=(1, 1, 1))     # outer faces only: not a slab of a 2-slab group
    with pytest.raises(ValueError):
        capi.SlabGroup([a, c])
