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
