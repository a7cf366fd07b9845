"""Deterministic synthetic initial grids for the BASELINE.json configs.

Plain numpy; the same arrays feed the reference oracle (as raw member-major files), the C
restatement and the CUDA path, so no generator has to be restated in three languages.
"""
import numpy as np

LBM_MEMBERS = ["C", "N", "E", "W", "S", "T", "B", "NW", "SW", "NE", "SE", "TW", "BW", "TE", "BE",
               "TN", "BN", "TS", "BS", "density", "velocityX", "velocityY", "velocityZ", "state"]
LBM_STATES = dict(LIQUID=0, WEST_NOSLIP=1, EAST_NOSLIP=2, TOP=3, BOTTOM=4, NORTH_ACC=5, SOUTH_NOSLIP=6)


def splitmix64(x):
    """Vectorised splitmix64 finaliser over uint64 numpy arrays."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(n, seed, offset=0):
    """n doubles in [0, 1): splitmix64(seed ^ index) / 2**64 (top 53 bits)."""
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    bits = splitmix64(idx ^ np.uint64(seed))
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def jacobi_grid(nx, ny, nz, seed=42, z0=0, nz_total=None):
    """Config 1/3: the hot cube of src/examples/jacobi3d/main.cpp:54-72 (value 0.99999999999,
    origin 5N/128, edge 50N/128) plus uniform noise. z0/nz_total select a slab of a taller
    global grid (multi-GPU weak scaling) with global indexing, so slabs tile seamlessly."""
    nz_total = nz if nz_total is None else nz_total
    plane = nx * ny
    v = uniform01(plane * nz, seed, offset=plane * z0).reshape(nz, ny, nx)
    off, size = nx * 5 // 128, nx * 50 // 128
    zs = np.arange(z0, z0 + nz)
    zmask = (zs >= off) & (zs < off + size)
    if size > 0 and zmask.any():
        sub = v[zmask]
        sub[:, off:off + size, off:off + size] = 0.99999999999
        v[zmask] = sub
    return v


def jacobi_box(nx, ny, nz, lo, hi, seed=42, zmap=None):
    """jacobi_grid(nx, ny, nz, seed)[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] (z, y, x) without building the whole grid —
    the input of a windowed oracle at full size. zmap: plane indices to take instead of range(lo[0], hi[0])."""
    zs = np.arange(lo[0], hi[0]) if zmap is None else np.asarray(zmap)
    ys, xs = np.arange(lo[1], hi[1]), np.arange(lo[2], hi[2])
    idx = ((zs[:, None, None].astype(np.uint64) * np.uint64(ny) + ys[None, :, None].astype(np.uint64)) * np.uint64(nx)
           + xs[None, None, :].astype(np.uint64))
    bits = splitmix64(idx ^ np.uint64(seed))
    v = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    off, size = nx * 5 // 128, nx * 50 // 128
    hot = ((zs >= off) & (zs < off + size))[:, None, None] & ((ys >= off) & (ys < off + size))[None, :, None] \
        & ((xs >= off) & (xs < off + size))[None, None, :]
    v[hot] = 0.99999999999
    return v


def gol_grid(nx, ny, seed=7, density=0.35):
    """Config 2: Bernoulli(density) soup plus the glider / Diehard / Acorn of
    src/examples/gameoflife/main.cpp:69-108."""
    g = (uniform01(nx * ny, seed) < density).astype(np.uint8).reshape(ny, nx)
    cells = [(11, 10), (12, 11), (10, 12), (11, 12), (12, 12),
             (55, 70), (56, 70), (56, 71), (60, 71), (61, 71), (62, 71), (61, 69),
             (111, 30), (113, 31), (110, 32), (111, 32), (113, 32), (114, 32), (115, 32)]
    for x, y in cells:
        if x < nx and y < ny:
            g[y, x] = 1
    return g


def lbm_states(nx, ny, nz, z0=0, nz_total=None):
    """Wall states by face as src/examples/latticeboltzmann/main.cpp:249-287 (later
    assignments win at edges)."""
    nz_total = nz if nz_total is None else nz_total
    s = np.zeros((nz, ny, nx), dtype=np.int32)
    s[:, :, 0] = LBM_STATES["WEST_NOSLIP"]
    s[:, :, nx - 1] = LBM_STATES["EAST_NOSLIP"]
    s[:, 0, :] = LBM_STATES["SOUTH_NOSLIP"]
    s[:, ny - 1, :] = LBM_STATES["NORTH_ACC"]
    zs = np.arange(z0, z0 + nz)
    s[zs == 0] = LBM_STATES["BOTTOM"]
    s[zs == nz_total - 1] = LBM_STATES["TOP"]
    return s


def lbm_states_box(nx, ny, nz_total, lo, hi):
    """lbm_states of the global nx x ny x nz_total cavity on the box [lo, hi) (z, y, x)"""
    zs, ys, xs = (np.arange(lo[i], hi[i]) for i in range(3))
    s = np.zeros((len(zs), len(ys), len(xs)), dtype=np.int32)
    s[:, :, xs == 0] = LBM_STATES["WEST_NOSLIP"]
    s[:, :, xs == nx - 1] = LBM_STATES["EAST_NOSLIP"]
    s[:, ys == 0, :] = LBM_STATES["SOUTH_NOSLIP"]
    s[:, ys == ny - 1, :] = LBM_STATES["NORTH_ACC"]
    s[zs == 0] = LBM_STATES["BOTTOM"]
    s[zs == nz_total - 1] = LBM_STATES["TOP"]
    return s


def lbm_grid(nx, ny, nz, seed=11, noise=0.0, z0=0, nz_total=None):
    """Config 4: lid-driven cavity; rest population C = 1 (rho = 1), the others 0 (optionally
    with a little noise so that every population path is exercised). Returns the raw
    member-major float32/int32 buffer as a (24, nz, ny, nx) float32 array (state bit-cast)."""
    cells = nx * ny * nz
    raw = np.zeros((24, nz, ny, nx), dtype=np.float32)
    raw[0] = 1.0
    if noise:
        plane = nx * ny
        for m in range(19):
            r = uniform01(cells, seed + 1000 * m, offset=plane * z0).reshape(nz, ny, nx)
            raw[m] += (noise * (r - 0.5)).astype(np.float32)
    raw[19] = 1.0
    raw[23] = lbm_states(nx, ny, nz, z0, nz_total).view(np.float32)
    return raw


def nbody_cells(ncx, ncy, ncz, cap=32, edge=2.5, spacing=1.058, jitter=0.1, vel=0.0, seed=1234, dtype=np.float32,
                z0=0):
    """Config 5: particles on a cubic lattice of `spacing` (density ~0.844) jittered by +-jitter,
    binned into BoxCell containers of edge `edge`; optional uniform random velocities in
    [-vel, vel] (tests use them to force particles across cell faces). z0 = first cell plane of a
    slab (global particle ids and positions). Returns (counts int32 [ncz][ncy][ncx],
    parts dtype [ncz][ncy][ncx][cap][6])."""
    lx, ly = ncx * edge, ncy * edge
    zlo, zhi = z0 * edge, (z0 + ncz) * edge
    npx, npy = int(lx / spacing), int(ly / spacing)
    k0, k1 = int(np.ceil(zlo / spacing - 0.5)), int(np.ceil(zhi / spacing - 0.5))
    ix, iy, iz = np.meshgrid(np.arange(npx), np.arange(npy), np.arange(k0, k1), indexing="ij")
    ids = ((iz.astype(np.uint64) * np.uint64(1 << 20) + iy.astype(np.uint64)) * np.uint64(1 << 20) + ix.astype(np.uint64)).reshape(-1)
    n = ids.size

    def rnd(salt):
        bits = splitmix64(ids ^ np.uint64(seed + salt))
        return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    pos = np.stack([(ix.reshape(-1) + 0.5) * spacing + jitter * (2 * rnd(1) - 1),
                    (iy.reshape(-1) + 0.5) * spacing + jitter * (2 * rnd(2) - 1),
                    (iz.reshape(-1) + 0.5) * spacing + jitter * (2 * rnd(3) - 1)], axis=1).astype(dtype)
    v = np.stack([vel * (2 * rnd(4) - 1), vel * (2 * rnd(5) - 1), vel * (2 * rnd(6) - 1)], axis=1).astype(dtype)
    # bin with the reference's position checker (origin <= pos < origin + edge, in double)
    pd = pos.astype(np.float64)
    c = np.floor(pd / edge).astype(np.int64)
    c[:, 2] -= z0
    inside = ((c >= 0) & (c < np.array([ncx, ncy, ncz]))).all(axis=1)
    pos, v, c = pos[inside], v[inside], c[inside]
    cell = (c[:, 2] * ncy + c[:, 1]) * ncx + c[:, 0]
    order = np.argsort(cell, kind="stable")
    cell, pos, v = cell[order], pos[order], v[order]
    counts = np.bincount(cell, minlength=ncx * ncy * ncz).astype(np.int32)
    if counts.max() > cap:
        raise ValueError("cell capacity %d exceeded (%d)" % (cap, counts.max()))
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    slot = np.arange(cell.size) - start[cell]
    parts = np.zeros((ncx * ncy * ncz, cap, 6), dtype=dtype)
    parts[cell, slot, 0:3] = pos
    parts[cell, slot, 3:6] = v
    return counts.reshape(ncz, ncy, ncx), parts.reshape(ncz, ncy, ncx, cap, 6)


def container_cells(nx, ny, nz=1, n_dims=3, torus=False, cap=16, maxnb=20, seed=77, edge=False, fill=0.75,
                    min_neighbors=1, chunk=1 << 15):
    """ContainerCell grids of ID-keyed mesh elements (oracle/models/container.h) in the interchange format of
    include/b200geo.h: container c holds counts[c] elements (about `fill` of the capacity; some containers are empty)
    with the ascending ids 1 + c * cap + slot, as the Voronoi example's initializer numbers them
    (src/examples/voronoi/main.cpp:153-156); every element lists min_neighbors..maxnb neighbour ids, in random
    order, drawn from its own container and the 3^n_dims - 1 around it (across the seam on a torus); with `edge` a
    lookup beyond a Cube boundary names an element of the edge container. Returns (box, edge_box or None): dicts of
    counts [nz][ny][nx], ids / values / influx / nb_counts [..][cap], nb_ids [..][cap][maxnb]."""
    assert n_dims in (2, 3) and (n_dims == 3 or nz == 1)
    ncont = nx * ny * nz
    cidx = np.arange(ncont, dtype=np.uint64)

    def h(x, salt):
        return splitmix64(x ^ np.uint64((seed * 1000003 + salt) & 0xFFFFFFFFFFFF))

    u = (h(cidx, 1) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    counts = np.minimum(cap, np.floor(u * (2 * fill * cap + 1))).astype(np.int32)
    counts[(h(cidx, 2) % np.uint64(16)) == 0] = 0          # some empty containers
    if ncont > 0 and counts.max() == 0:
        counts[0] = 1
    n_edge = max(1, cap // 2) if edge else 0
    slot = np.arange(cap, dtype=np.int64)
    ids = (1 + np.arange(ncont, dtype=np.int64)[:, None] * cap + slot[None, :]).astype(np.int32)
    gslot = np.arange(ncont * cap, dtype=np.uint64).reshape(ncont, cap)
    values = (h(gslot, 3) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    influx = np.where(h(gslot, 4) % np.uint64(8) == 0, (h(gslot, 5) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0 * 0.125, 0.0)
    nb_counts = (min_neighbors + h(gslot, 6) % np.uint64(maxnb - min_neighbors + 1)).astype(np.int32)
    live = slot[None, :] < counts[:, None]
    ids[~live] = 0
    values[~live] = 0
    influx[~live] = 0
    nb_counts[~live] = 0
    nb_ids = np.zeros((ncont, cap, maxnb), dtype=np.int32)
    dims = (nx, ny, nz)
    cnt64 = counts.astype(np.int64)
    jj = np.arange(maxnb, dtype=np.int32)[None, None, :]
    for c0 in range(0, ncont, chunk):
        c1 = min(ncont, c0 + chunk)
        c = np.arange(c0, c1, dtype=np.int64)
        r = h(np.arange(c0 * cap * maxnb, c1 * cap * maxnb, dtype=np.uint64), 7).reshape(c1 - c0, cap, maxnb)
        coord = (c % nx, (c // nx) % ny, c // (nx * ny))
        t = np.zeros(r.shape, dtype=np.int64)
        outside = np.zeros(r.shape, dtype=bool)
        stride = 1
        for a in range(n_dims):
            p = coord[a][:, None, None] + ((r >> np.uint64(8 * a)) % np.uint64(3)).astype(np.int64) - 1
            if torus:
                p = (p + dims[a]) % dims[a]
            else:
                outside |= (p < 0) | (p >= dims[a])
                np.clip(p, 0, dims[a] - 1, out=p)
            t += p * stride
            stride *= dims[a]
        own = c[:, None, None]
        # lookups that left a Cube grid without an edge container, or that hit an empty container, fall back on the own one
        if not edge:
            t = np.where(outside, own, t)
        t = np.where(cnt64[t] == 0, own, t)
        pick = (r >> np.uint64(24)).astype(np.int64)
        target = 1 + t * cap + pick % np.maximum(cnt64[t], 1)
        if edge:
            target = np.where(outside, 1 + ncont * cap + pick % n_edge, target)
        nb_ids[c0:c1] = np.where(jj < nb_counts[c0:c1, :, None], target, 0)
    shape = (nz, ny, nx) if n_dims == 3 else (ny, nx)
    box = {"counts": counts.reshape(shape), "ids": ids.reshape(shape + (cap,)), "values": values.reshape(shape + (cap,)),
           "influx": influx.reshape(shape + (cap,)), "nb_counts": nb_counts.reshape(shape + (cap,)),
           "nb_ids": nb_ids.reshape(shape + (cap, maxnb))}
    edge_box = None
    if edge:
        es = np.arange(cap, dtype=np.uint64)
        edge_box = {"counts": np.array([n_edge], dtype=np.int32),
                    "ids": np.where(np.arange(cap) < n_edge, 1 + ncont * cap + np.arange(cap), 0).astype(np.int32),
                    "values": np.where(np.arange(cap) < n_edge, 2.0 + (h(es, 8) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0, 0.0),
                    "influx": np.zeros(cap, dtype=np.float64), "nb_counts": np.zeros(cap, dtype=np.int32),
                    "nb_ids": np.zeros((cap, maxnb), dtype=np.int32)}
    return box, edge_box


def container_duplicate_ids(box):
    """Renumbers the elements of a container_cells() grid IN PLACE so that the same id occurs in several containers
    (id -> 1 + (id - 1) % (7 * cap)): which element hood[id] returns then depends on the search order of
    NeighborhoodAdapter::operator[] (own container first, then CoordBox order; storage/neighborhoodadapter.h:45-65).
    Ids stay ascending inside a container."""
    cap = box["ids"].shape[-1]
    m = 7 * cap
    for name in ("ids", "nb_ids"):
        a = box[name]
        a[a > 0] = 1 + (a[a > 0] - 1) % m
    return box
