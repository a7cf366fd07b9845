"""TEST INFRASTRUCTURE — ctypes access to oracle/liboracle.so (the C restatement) and a runner
for the reference-built binaries under oracle/_ref/. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module."""
import ctypes
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def jacobi(kind, torus, grid, steps, edge=0.0):
    g = np.ascontiguousarray(grid, dtype=np.float64)
    nz, ny, nx = g.shape
    out = np.empty_like(g)
    rc = lib().oracle_jacobi(kind, int(torus), nx, ny, nz, steps, ctypes.c_double(edge), _p(g), _p(out))
    assert rc == 0, rc
    return out


def gol(torus, grid, steps, edge_alive=0):
    g = np.ascontiguousarray(grid, dtype=np.uint8)
    ny, nx = g.shape
    out = np.empty_like(g)
    rc = lib().oracle_gol(int(torus), nx, ny, steps, int(edge_alive), _p(g), _p(out))
    assert rc == 0, rc
    return out


def lbm(raw, steps):
    g = np.ascontiguousarray(raw, dtype=np.float32)
    m, nz, ny, nx = g.shape
    assert m == 24
    out = np.empty_like(g)
    rc = lib().oracle_lbm(nx, ny, nz, steps, _p(g), _p(out))
    assert rc == 0, rc
    return out


def nbody(counts, parts, steps, dt=0.005, cutoff=2.5, edge=2.5, origin=(0, 0, 0)):
    """counts int32 [nz][ny][nx], parts REAL [nz][ny][nx][cap][6] -> (counts, parts) after `steps`.
    Raises IndexError when a container overflows (std::out_of_range in the reference)."""
    c = np.ascontiguousarray(counts, dtype=np.int32)
    p = np.ascontiguousarray(parts)
    assert p.dtype in (np.float32, np.float64) and p.shape[:3] == c.shape and p.shape[4] == 6
    nz, ny, nx = c.shape
    co, po = np.empty_like(c), np.empty_like(p)
    f = lib().oracle_nbody_at
    f.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double] * 3 + [ctypes.c_void_p] * 5
    org = (ctypes.c_int * 3)(*[int(v) for v in origin])
    rc = f(p.dtype.itemsize, nx, ny, nz, p.shape[3], steps, dt, cutoff, edge, org, _p(c), _p(p), _p(co), _p(po))
    if rc == -3:
        raise IndexError("capacity exceeded")
    assert rc == 0, rc
    return co, po


def run_ref_nbody(counts, parts, steps, dt=0.005, cutoff=2.5, edge=2.5, omp=False, threads=None, want_output=True):
    """The reference's own SerialSimulator/OpenMPSimulator over BoxCell containers (oracle/_ref/lgd_ref_nbody,
    capacity 32). Returns ((counts, parts) or None, stats)."""
    c = np.ascontiguousarray(counts, dtype=np.int32)
    p = np.ascontiguousarray(parts)
    assert p.shape[3] == 32, "the reference binary is built for FixedArray capacity 32"
    nz, ny, nx = c.shape
    env = dict(os.environ)
    if omp:
        env["OMP_PROC_BIND"] = "true"
        env["OMP_PLACES"] = "cores"
        if threads:
            env["OMP_NUM_THREADS"] = str(threads)
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.raw"), os.path.join(tmp, "out.raw")
        with open(fin, "wb") as f:
            f.write(c.tobytes())
            f.write(p.tobytes())
        cmd = [ref_binary("nbody"), "nbody", str(nx), str(ny), str(nz), str(steps), fin, fout if want_output else "-",
               "--dt", repr(float(dt)), "--cutoff", repr(float(cutoff)), "--edge", repr(float(edge))]
        if p.dtype == np.float64:
            cmd.append("--double")
        if omp:
            cmd.append("--omp")
        res = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if res.returncode == 4 and "capacity exceeded" in res.stderr:
            raise IndexError("capacity exceeded")
        if res.returncode != 0:
            raise RuntimeError(res.stderr[-2000:])
        stats = json.loads(res.stdout.strip().splitlines()[-1])
        out = None
        if want_output:
            raw = np.fromfile(fout, dtype=np.uint8)
            co = raw[:c.nbytes].view(np.int32).reshape(c.shape)
            po = raw[c.nbytes:].view(p.dtype).reshape(p.shape)
            out = (co, po)
    return out, stats


CONTAINER_FIELDS = ("counts", "ids", "values", "influx", "nb_counts", "nb_ids")
_CONTAINER_TYPES = (np.int32, np.int32, np.float64, np.float64, np.int32, np.int32)


def _container_arrays(box):
    return [np.ascontiguousarray(box[n], dtype=t) for n, t in zip(CONTAINER_FIELDS, _CONTAINER_TYPES)]


def container(box, steps, n_dims=3, torus=False, edge=None):
    """box: dict of the six arrays of a whole grid (include/b200geo.h; counts [nz][ny][nx] or [ny][nx]); edge: the
    same for the one edge container or None. Returns the temperatures [..][cap] after `steps`. Raises KeyError(id)
    for an id that is not in the neighbourhood (std::logic_error "id not found" in the reference)."""
    a = _container_arrays(box)
    shape = a[0].shape
    assert len(shape) == n_dims
    nz, ny, nx = (shape if n_dims == 3 else (1,) + shape)
    cap, maxnb = a[5].shape[-2], a[5].shape[-1]
    out = np.empty_like(a[2])
    missing = ctypes.c_int32(0)
    e = _container_arrays(edge) if edge is not None else None
    f = lib().oracle_container
    f.argtypes = [ctypes.c_int] * 8 + [ctypes.c_void_p] * 10 + [ctypes.POINTER(ctypes.c_int32)]
    rc = f(n_dims, int(torus), nx, ny, nz, cap, maxnb, steps, *[_p(v) for v in a],
           _p(e[0]) if e else None, _p(e[1]) if e else None, _p(e[2]) if e else None, _p(out), ctypes.byref(missing))
    if rc == -2:
        raise KeyError(int(missing.value))
    assert rc == 0, rc
    return out


def run_ref_container(box, steps, n_dims=3, torus=False, edge=None, omp=False, threads=None, want_output=True):
    """The reference's own SerialSimulator / OpenMPSimulator over ContainerCell<MeshElement, 16> containers with
    FixedArray<int, 20> neighbour lists (oracle/_ref/lgd_ref_container). Returns (dict of the six arrays or None, stats)."""
    a = _container_arrays(box)
    shape = a[0].shape
    nz, ny, nx = (shape if n_dims == 3 else (1,) + shape)
    assert a[5].shape[-2:] == (16, 20), "the reference binary is built for capacity 16 and 20 neighbour ids"
    env = dict(os.environ)
    if omp:
        env["OMP_PROC_BIND"] = "true"
        env["OMP_PLACES"] = "cores"
        if threads:
            env["OMP_NUM_THREADS"] = str(threads)
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.raw"), os.path.join(tmp, "out.raw")
        with open(fin, "wb") as f:
            for v in a:
                f.write(v.tobytes())
            if edge is not None:
                for v in _container_arrays(edge):
                    f.write(v.tobytes())
        cmd = [ref_binary("container"), "container", str(nx), str(ny), str(nz), str(steps), fin, fout if want_output else "-",
               "--dims", str(n_dims)]
        cmd += ["--torus"] if torus else []
        cmd += ["--edge"] if edge is not None else []
        cmd += ["--omp"] if omp else []
        res = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if res.returncode == 4 and "id not found" in res.stderr:
            raise KeyError(res.stderr.strip())
        if res.returncode != 0:
            raise RuntimeError(res.stderr[-2000:])
        stats = json.loads(res.stdout.strip().splitlines()[-1])
        out = None
        if want_output:
            raw = np.fromfile(fout, dtype=np.uint8)
            out, pos = {}, 0
            for n, v in zip(CONTAINER_FIELDS, a):
                out[n] = raw[pos:pos + v.nbytes].view(v.dtype).reshape(v.shape)
                pos += v.nbytes
    return out, stats


def _region(fn, grid_raw, dims, member_bytes, streaks, buf):
    nx, ny, nz = dims
    mb = np.asarray(member_bytes, dtype=np.int32)
    st = np.ascontiguousarray(streaks, dtype=np.int32).reshape(-1, 4)
    rc = fn(nx, ny, nz, len(mb), _p(mb), _p(grid_raw), _p(st), len(st), _p(buf))
    assert rc == 0, rc


def save_region(grid_raw, dims, member_bytes, streaks):
    st = np.asarray(streaks, dtype=np.int32).reshape(-1, 4)
    count = int((st[:, 3] - st[:, 0]).sum())
    buf = np.zeros(count * int(sum(member_bytes)), dtype=np.uint8)
    _region(lib().oracle_save_region, grid_raw, dims, member_bytes, st, buf)
    return buf


def load_region(grid_raw, dims, member_bytes, streaks, buf):
    _region(lib().oracle_load_region, grid_raw, dims, member_bytes, streaks, np.ascontiguousarray(buf))


# ------------------------------------------------------------------ reference binaries

def ref_binary(model):
    return os.path.join(HERE, "_ref", "lgd_ref_" + model)


def have_ref(model):
    return os.access(ref_binary(model), os.X_OK)


def run_ref(model, raw_in, dims, steps, omp=False, edge=None, threads=None, want_output=True, tile=0):
    """Run the reference's own SerialSimulator/OpenMPSimulator on raw_in (any numpy array whose
    bytes are the member-major grid); returns (raw_out_bytes as uint8 array, stats dict).
    tile > 0: raw_in holds only `tile` planes (nx x ny x tile) and is repeated along z."""
    nx, ny, nz = dims
    env = dict(os.environ)
    if omp:
        env["OMP_PROC_BIND"] = "true"
        env["OMP_PLACES"] = "cores"
        if threads:
            env["OMP_NUM_THREADS"] = str(threads)
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.raw"), os.path.join(tmp, "out.raw")
        np.ascontiguousarray(raw_in).tofile(fin)
        cmd = [ref_binary(model), model, str(nx), str(ny), str(nz), str(steps), fin,
               fout if want_output else "-"]
        if omp:
            cmd.append("--omp")
        if edge is not None:
            cmd += ["--edge", repr(float(edge))]
        if tile:
            cmd += ["--tile", str(int(tile))]
        res = subprocess.run(cmd, env=env, check=True, capture_output=True, text=True)
        stats = json.loads(res.stdout.strip().splitlines()[-1])
        out = np.fromfile(fout, dtype=np.uint8) if want_output else None
    return out, stats
