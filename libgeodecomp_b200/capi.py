"""ctypes binding of the C ABI in include/b200geo.h (libb200geo.so).

This is the Python twin of the C++ façade in include/libgeodecomp_b200/: it only marshals
arguments and maps status codes to exceptions the way the façade maps them to the reference's
C++ exceptions (std::invalid_argument -> ValueError, std::logic_error -> RuntimeError subclass
LogicError, std::out_of_range -> IndexError, "CUDA error" -> CudaError). There is no fallback:
if the shared library is missing, importing the engine fails loudly.
"""
import ctypes
import os

import numpy as np

MAX_MEMBERS = 32
HOST, CUDA_DEVICE = 0, 1
GHOST_EDGE, GHOST_WRAP, GHOST_PEER = 0, 1, 2
KERNEL_JACOBI6, KERNEL_JACOBI7, KERNEL_JACOBI27, KERNEL_GOL, KERNEL_LBM_D3Q19, KERNEL_NBODY = 1, 2, 3, 4, 5, 6
KERNEL_CONTAINER = 7

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200geo.so")


class LogicError(RuntimeError):
    """std::logic_error"""


class CudaError(RuntimeError):
    """std::runtime_error("CUDA error"), misc/cudautil.h:48-55"""


class GridDesc(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32 * 3),
                ("ghost", ctypes.c_int32 * 3),
                ("ghost_mode", (ctypes.c_int32 * 2) * 3),
                ("n_members", ctypes.c_int32),
                ("member_bytes", ctypes.c_int32 * MAX_MEMBERS)]


class BoxGridDesc(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32 * 3),
                ("ghost_mode", (ctypes.c_int32 * 2) * 3),
                ("capacity", ctypes.c_int32),
                ("real_bytes", ctypes.c_int32),
                ("cell_origin", ctypes.c_int32 * 3),
                ("cell_edge", ctypes.c_double)]


class ContainerGridDesc(ctypes.Structure):
    _fields_ = [("n_dims", ctypes.c_int32),
                ("dim", ctypes.c_int32 * 3),
                ("ghost_mode", (ctypes.c_int32 * 2) * 3),
                ("capacity", ctypes.c_int32),
                ("max_neighbors", ctypes.c_int32)]


class ContainerBox(ctypes.Structure):
    """b200geo_container_box: the six arrays of a box of containers (interchange format of include/b200geo.h)"""
    FIELDS = ("counts", "ids", "values", "influx", "nb_counts", "nb_ids")
    _fields_ = [(n, ctypes.c_void_p) for n in FIELDS]


class NBodyParams(ctypes.Structure):
    _fields_ = [("dt", ctypes.c_double), ("cutoff", ctypes.c_double), ("nano_steps", ctypes.c_int32)]


_lib = None

# every symbol include/b200geo.h declares: (name, restype, argtypes)
_vp, _i32p, _i64p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64)
SYMBOLS = [
    ("b200geo_version", ctypes.c_char_p, []),
    ("b200geo_last_error", ctypes.c_char_p, []),
    ("b200geo_device_count", ctypes.c_int, []),
    ("b200geo_launch_count", ctypes.c_uint64, []),
    ("b200geo_set_tuning", ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    ("b200geo_grid_create", ctypes.c_int, [ctypes.POINTER(GridDesc), ctypes.c_int, ctypes.POINTER(_vp)]),
    ("b200geo_grid_create_uniform", ctypes.c_int, [ctypes.POINTER(GridDesc), ctypes.c_int, ctypes.c_int64, ctypes.POINTER(_vp)]),
    ("b200geo_grid_uniform_min_stride", ctypes.c_int, [ctypes.POINTER(GridDesc), _i64p]),
    ("b200geo_grid_plan", ctypes.c_int, [ctypes.POINTER(GridDesc), ctypes.c_int64, _i64p, _i64p]),
    ("b200geo_grid_member_stride", ctypes.c_int, [_vp, _i64p]),
    ("b200geo_grid_destroy", ctypes.c_int, [_vp]),
    ("b200geo_grid_buffer_bytes", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64)]),
    ("b200geo_grid_device", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int)]),
    ("b200geo_grid_layout", ctypes.c_int, [_vp, ctypes.c_int, _i64p, _i64p, _i64p]),
    ("b200geo_grid_member_ptr", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]),
    ("b200geo_grid_set_edge", ctypes.c_int, [_vp, _vp, _vp]),
    ("b200geo_grid_get_edge", ctypes.c_int, [_vp, _vp]),
    ("b200geo_grid_load_member", ctypes.c_int, [_vp, ctypes.c_int, _i32p, _i32p, _vp, ctypes.c_int, ctypes.c_int, _vp]),
    ("b200geo_grid_save_member", ctypes.c_int, [_vp, ctypes.c_int, _i32p, _i32p, _vp, ctypes.c_int, _vp]),
    ("b200geo_grid_load_region", ctypes.c_int, [_vp, _i32p, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp]),
    ("b200geo_grid_save_region", ctypes.c_int, [_vp, _i32p, ctypes.c_int, _vp, ctypes.c_int, _vp]),
    ("b200geo_step", ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_uint32, ctypes.c_uint32, _vp]),
    ("b200geo_update_box", ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_uint32, _i32p, _i32p, _vp]),
    ("b200geo_update_box_n", ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_uint32, _i32p, _i32p, ctypes.c_uint32, _vp]),
    ("b200geo_swap", ctypes.c_int, [_vp]),
    ("b200geo_refresh_ghosts", ctypes.c_int, [_vp, _vp]),
    ("b200geo_sync", ctypes.c_int, [_vp]),
    ("b200geo_grid_sync", ctypes.c_int, [_vp, _vp]),
    ("b200geo_stream_create", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    ("b200geo_stream_destroy", ctypes.c_int, [ctypes.c_int, _vp]),
    ("b200geo_stream_wait", ctypes.c_int, [ctypes.c_int, _vp, _vp]),
    ("b200geo_device_alloc", ctypes.c_int, [ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p)]),
    ("b200geo_device_free", ctypes.c_int, [ctypes.c_int, _vp]),
    ("b200geo_device_copy", ctypes.c_int, [ctypes.c_int, _vp, ctypes.c_int, _vp, ctypes.c_uint64]),
    ("b200geo_host_alloc", ctypes.c_int, [ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p)]),
    ("b200geo_host_free", ctypes.c_int, [_vp]),
    ("b200geo_halo_block", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_uint64)]),
    ("b200geo_halo_block_in", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_uint64)]),
    ("b200geo_grid_ipc_export", ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    ("b200geo_grid_ipc_open", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    ("b200geo_halo_push", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    ("b200geo_halo_mark_valid", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    ("b200geo_group_create", ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]),
    ("b200geo_group_destroy", ctypes.c_int, [_vp]),
    ("b200geo_group_invalidate", ctypes.c_int, [_vp]),
    ("b200geo_group_exchange", ctypes.c_int, [_vp]),
    ("b200geo_group_step", ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_uint32, ctypes.c_uint32]),
    ("b200geo_group_step_with", ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint32]),
    ("b200geo_group_sync", ctypes.c_int, [_vp]),
    ("b200geo_group_stats", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64)]),
    ("b200geo_boxgrid_create", ctypes.c_int, [ctypes.POINTER(BoxGridDesc), ctypes.c_int, ctypes.POINTER(_vp)]),
    ("b200geo_boxgrid_destroy", ctypes.c_int, [_vp]),
    ("b200geo_boxgrid_load", ctypes.c_int, [_vp, _i32p, _i32p, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp]),
    ("b200geo_boxgrid_save", ctypes.c_int, [_vp, _i32p, _i32p, _vp, _vp, ctypes.c_int, _vp]),
    ("b200geo_boxgrid_step", ctypes.c_int, [_vp, ctypes.POINTER(NBodyParams), ctypes.c_uint32, ctypes.c_uint32, _vp]),
    ("b200geo_boxgrid_check", ctypes.c_int, [_vp, _vp]),
    ("b200geo_boxgrid_halo_block", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_uint64)]),
    ("b200geo_boxgrid_halo_mark_valid", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    ("b200geo_boxgroup_create", ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.POINTER(_vp)]),
    ("b200geo_boxgroup_destroy", ctypes.c_int, [_vp]),
    ("b200geo_boxgroup_step", ctypes.c_int, [_vp, ctypes.POINTER(NBodyParams), ctypes.c_uint32, ctypes.c_uint32]),
    ("b200geo_boxgroup_sync", ctypes.c_int, [_vp]),
    ("b200geo_boxgroup_stats", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64)]),
    ("b200geo_containergrid_create", ctypes.c_int, [ctypes.POINTER(ContainerGridDesc), ctypes.c_int, ctypes.POINTER(_vp)]),
    ("b200geo_containergrid_destroy", ctypes.c_int, [_vp]),
    ("b200geo_containergrid_load", ctypes.c_int, [_vp, _i32p, _i32p, ctypes.POINTER(ContainerBox), ctypes.c_int, _vp]),
    ("b200geo_containergrid_save", ctypes.c_int, [_vp, _i32p, _i32p, ctypes.POINTER(ContainerBox), ctypes.c_int, _vp]),
    ("b200geo_containergrid_set_edge", ctypes.c_int, [_vp, ctypes.POINTER(ContainerBox)]),
    ("b200geo_containergrid_get_edge", ctypes.c_int, [_vp, ctypes.POINTER(ContainerBox)]),
    ("b200geo_containergrid_step", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, _vp]),
    ("b200geo_containergrid_stats", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64)]),
    ("b200geo_stats_enable", ctypes.c_int, [_vp, ctypes.c_int]),
    ("b200geo_stats", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_double)]),
]


def lib_path():
    return _LIB_PATH


def lib():
    """Load libb200geo.so; raises ImportError (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                "libb200geo.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C libgeodecomp_b200/csrc` (there is no CPU fallback)")
        handle = ctypes.CDLL(_LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc):
    if rc >= 0:
        return rc
    msg = lib().b200geo_last_error().decode()
    if rc == -1:
        raise ValueError(msg)
    if rc == -2:
        raise LogicError(msg)
    if rc == -3:
        raise IndexError(msg)
    if rc == -5:
        raise MemoryError(msg)
    raise CudaError(msg)


def _i3(v):
    return (ctypes.c_int32 * 3)(*[int(x) for x in v])


def _ptr(buf):
    """address of a numpy array, a torch tensor, an int address or None"""
    if buf is None:
        return None
    if isinstance(buf, int):
        return ctypes.c_void_p(buf)
    if isinstance(buf, np.ndarray):
        return ctypes.c_void_p(buf.ctypes.data)
    if hasattr(buf, "data_ptr"):
        return ctypes.c_void_p(buf.data_ptr())
    raise TypeError("unsupported buffer type %r" % type(buf))


class DeviceBlock:
    """A contiguous byte range of device memory, exportable to torch (for NCCL send/recv) through
    __cuda_array_interface__."""

    def __init__(self, ptr, nbytes, owner):
        self.ptr, self.nbytes, self.owner = ptr, nbytes, owner
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}

    def as_tensor(self):
        """zero-copy view on the device the block lives on (NOT torch's current device: torch would copy the block
        there and NCCL would send and receive on the copy)"""
        import torch
        dev = getattr(self.owner, "device", None)
        t = torch.as_tensor(self, device=torch.device("cuda", int(dev)) if dev is not None else "cuda")
        if t.data_ptr() != self.ptr:
            raise LogicError("torch copied the halo block instead of viewing it (device mismatch)")
        return t


class DeviceGrid:
    """Thin object wrapper around a b200geo_grid handle."""

    def __init__(self, dim, member_bytes, ghost=(1, 1, 1), ghost_mode=None, device=0, member_stride=None):
        """member_stride: None = the default padded layout; an element count (or 0 = the smallest valid one) = the
        uniform element layout of b200geo_grid_create_uniform."""
        self._h = None
        desc = GridDesc()
        dim = list(dim) + [1] * (3 - len(dim))
        ghost = list(ghost) + [0] * (3 - len(ghost))
        for i in range(3):
            desc.dim[i] = int(dim[i])
            desc.ghost[i] = int(ghost[i])
            for s in range(2):
                desc.ghost_mode[i][s] = int(ghost_mode[i][s]) if ghost_mode is not None else GHOST_EDGE
        if len(member_bytes) > MAX_MEMBERS:
            raise ValueError("too many members")
        desc.n_members = len(member_bytes)
        for m, b in enumerate(member_bytes):
            desc.member_bytes[m] = int(b)
        h = ctypes.c_void_p()
        if member_stride is None:
            check(lib().b200geo_grid_create(ctypes.byref(desc), int(device), ctypes.byref(h)))
        else:
            least = ctypes.c_int64()
            check(lib().b200geo_grid_uniform_min_stride(ctypes.byref(desc), ctypes.byref(least)))
            self.member_stride = int(member_stride) if member_stride else least.value
            check(lib().b200geo_grid_create_uniform(ctypes.byref(desc), int(device), self.member_stride, ctypes.byref(h)))
        self._h = h
        self.dim, self.ghost, self.member_bytes, self.device = tuple(dim), tuple(ghost), list(member_bytes), device
        self.cell_bytes = int(sum(member_bytes))

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.b200geo_grid_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- queries
    def buffer_bytes(self):
        v = ctypes.c_uint64()
        check(lib().b200geo_grid_buffer_bytes(self._h, ctypes.byref(v)))
        return v.value

    def layout(self, member=0):
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(lib().b200geo_grid_layout(self._h, member, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def member_ptr(self, member, which=0):
        p = ctypes.c_void_p()
        check(lib().b200geo_grid_member_ptr(self._h, member, which, ctypes.byref(p)))
        return p.value

    # -- edge cell
    def set_edge(self, cell_bytes, stream=None):
        buf = np.frombuffer(bytes(cell_bytes), dtype=np.uint8)
        if buf.size != self.cell_bytes:
            raise ValueError("edge cell must be %d bytes" % self.cell_bytes)
        check(lib().b200geo_grid_set_edge(self._h, _ptr(buf), stream))

    def get_edge(self):
        buf = np.zeros(self.cell_bytes, dtype=np.uint8)
        check(lib().b200geo_grid_get_edge(self._h, _ptr(buf)))
        return buf.tobytes()

    # -- bulk I/O
    def load_member(self, member, src, origin=(0, 0, 0), dim=None, location=HOST, both=True, stream=None):
        dim = self.dim if dim is None else dim
        check(lib().b200geo_grid_load_member(self._h, member, _i3(origin), _i3(dim), _ptr(src), location,
                                             1 if both else 0, stream))

    def save_member(self, member, dst, origin=(0, 0, 0), dim=None, location=HOST, stream=None):
        dim = self.dim if dim is None else dim
        check(lib().b200geo_grid_save_member(self._h, member, _i3(origin), _i3(dim), _ptr(dst), location, stream))

    def load_region(self, streaks, buf, location=HOST, both=True, stream=None):
        st = np.ascontiguousarray(streaks, dtype=np.int32).reshape(-1, 4)
        check(lib().b200geo_grid_load_region(self._h, st.ctypes.data_as(_i32p), len(st), _ptr(buf), location,
                                             1 if both else 0, stream))

    def save_region(self, streaks, buf, location=HOST, stream=None):
        st = np.ascontiguousarray(streaks, dtype=np.int32).reshape(-1, 4)
        check(lib().b200geo_grid_save_region(self._h, st.ctypes.data_as(_i32p), len(st), _ptr(buf), location, stream))

    # -- hot path
    def step(self, kernel, n_steps=1, first_nano_step=0, params=None, stream=None):
        check(lib().b200geo_step(self._h, kernel, _ptr(params), first_nano_step, n_steps, stream))

    def update_box(self, kernel, origin, dim, nano_step=0, params=None, stream=None, n_sweeps=1):
        check(lib().b200geo_update_box_n(self._h, kernel, _ptr(params), nano_step, _i3(origin), _i3(dim),
                                         int(n_sweeps), stream))

    def swap(self):
        check(lib().b200geo_swap(self._h))

    def refresh_ghosts(self, stream=None):
        check(lib().b200geo_refresh_ghosts(self._h, stream))

    # -- halo
    def halo_block(self, member, side, kind, width=1, which=0):
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        check(lib().b200geo_halo_block_in(self._h, member, side, kind, width, which, ctypes.byref(p), ctypes.byref(n)))
        return DeviceBlock(p.value, n.value, self)

    def halo_mark_valid(self, side, width):
        check(lib().b200geo_halo_mark_valid(self._h, side, width))

    def ipc_export(self, which):
        buf = ctypes.create_string_buffer(64)
        check(lib().b200geo_grid_ipc_export(self._h, which, buf))
        return buf.raw

    def ipc_open(self, side, which, handle):
        check(lib().b200geo_grid_ipc_open(self._h, side, which, ctypes.create_string_buffer(handle, 64)))

    def halo_push(self, side, width=1, stream=None):
        check(lib().b200geo_halo_push(self._h, side, width, stream))

    # -- statistics
    def stats_enable(self, on=True):
        check(lib().b200geo_stats_enable(self._h, 1 if on else 0))

    def stats(self):
        out = (ctypes.c_double * 3)()
        check(lib().b200geo_stats(self._h, out))
        return {"update_s": out[0], "ghost_s": out[1], "sweeps": int(out[2])}


class SlabGroup:
    """b200geo_group: the slabs of one simulation space on several GPUs of this box, driven by one
    host thread (rim-first schedule, direct NVLink copies between neighbouring slabs)."""

    def __init__(self, grids, periodic=False):
        self._h = None
        self.grids = list(grids)
        arr = (ctypes.c_void_p * len(self.grids))(*[g._h for g in self.grids])
        h = ctypes.c_void_p()
        check(lib().b200geo_group_create(arr, len(self.grids), 1 if periodic else 0, ctypes.byref(h)))
        self._h = h

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.b200geo_group_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def invalidate(self):
        check(lib().b200geo_group_invalidate(self._h))

    def exchange(self):
        check(lib().b200geo_group_exchange(self._h))

    def step(self, kernel, n_steps=1, first_nano_step=0, params=None):
        check(lib().b200geo_group_step(self._h, kernel, _ptr(params), first_nano_step, n_steps))

    def sync(self):
        check(lib().b200geo_group_sync(self._h))

    def stats(self):
        out = (ctypes.c_uint64 * 2)()
        check(lib().b200geo_group_stats(self._h, out))
        return {"exchanges": int(out[0]), "bytes": int(out[1])}


class BoxSlabGroup:
    """b200geo_boxgroup: slabs of BoxCell containers on several GPUs of this box, one host thread"""

    def __init__(self, grids):
        self._h = None
        self.grids = list(grids)
        arr = (ctypes.c_void_p * len(self.grids))(*[g._h for g in self.grids])
        h = ctypes.c_void_p()
        check(lib().b200geo_boxgroup_create(arr, len(self.grids), ctypes.byref(h)))
        self._h = h

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.b200geo_boxgroup_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, params, n_steps=1, first_nano_step=0):
        check(lib().b200geo_boxgroup_step(self._h, ctypes.byref(params), first_nano_step, n_steps))

    def sync(self):
        check(lib().b200geo_boxgroup_sync(self._h))

    def stats(self):
        out = (ctypes.c_uint64 * 2)()
        check(lib().b200geo_boxgroup_stats(self._h, out))
        return {"exchanges": int(out[0]), "bytes": int(out[1])}


class DeviceBoxGrid:
    """Thin object wrapper around a b200geo_boxgrid handle (BoxCell container grid, n-body)."""

    def __init__(self, dim, capacity, real_bytes, cell_edge, cell_origin=(0, 0, 0), ghost_mode=None, device=0):
        self._h = None
        desc = BoxGridDesc()
        for i in range(3):
            desc.dim[i] = int(dim[i])
            desc.cell_origin[i] = int(cell_origin[i])
            for s in range(2):
                desc.ghost_mode[i][s] = int(ghost_mode[i][s]) if ghost_mode is not None else GHOST_EDGE
        desc.capacity, desc.real_bytes, desc.cell_edge = int(capacity), int(real_bytes), float(cell_edge)
        h = ctypes.c_void_p()
        check(lib().b200geo_boxgrid_create(ctypes.byref(desc), int(device), ctypes.byref(h)))
        self._h = h
        self.dim, self.capacity, self.real_bytes, self.device = tuple(dim), int(capacity), int(real_bytes), device

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.b200geo_boxgrid_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, counts, particles, origin=(0, 0, 0), dim=None, location=HOST, both=True, stream=None):
        dim = self.dim if dim is None else dim
        check(lib().b200geo_boxgrid_load(self._h, _i3(origin), _i3(dim), _ptr(counts), _ptr(particles), location,
                                         1 if both else 0, stream))

    def save(self, counts, particles, origin=(0, 0, 0), dim=None, location=HOST, stream=None):
        dim = self.dim if dim is None else dim
        check(lib().b200geo_boxgrid_save(self._h, _i3(origin), _i3(dim), _ptr(counts), _ptr(particles), location, stream))

    def step(self, kernel, n_steps=1, first_nano_step=0, params=None, stream=None):
        if kernel != KERNEL_NBODY or not isinstance(params, NBodyParams):
            raise LogicError("a BoxCell grid steps with KERNEL_NBODY and NBodyParams")
        check(lib().b200geo_boxgrid_step(self._h, ctypes.byref(params), first_nano_step, n_steps, stream))

    def check(self, stream=None):
        """raises IndexError (std::out_of_range "capacity exceeded") if a container overflowed"""
        check(lib().b200geo_boxgrid_check(self._h, stream))

    def halo_block(self, member, side, kind, width=1, which=0):
        if width != 1:
            raise ValueError("BoxCell grids have a ghost zone of one container")
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        check(lib().b200geo_boxgrid_halo_block(self._h, member, side, kind, which, ctypes.byref(p), ctypes.byref(n)))
        return DeviceBlock(p.value, n.value, self)

    def halo_mark_valid(self, side, width):
        check(lib().b200geo_boxgrid_halo_mark_valid(self._h, side, width))

    def stats_enable(self, on=True):
        pass

    def stats(self):
        return {}


def container_box(arrays):
    """ContainerBox from a dict / sequence of the six arrays (None entries = NULL: skipped by save)"""
    if isinstance(arrays, dict):
        arrays = [arrays.get(n) for n in ContainerBox.FIELDS]
    box = ContainerBox()
    for name, a in zip(ContainerBox.FIELDS, arrays):
        p = _ptr(a)
        setattr(box, name, p.value if p is not None else None)
    return box


class DeviceContainerGrid:
    """Thin object wrapper around a b200geo_containergrid handle (ContainerCell grid of ID-keyed cargo)."""

    def __init__(self, dim, capacity, max_neighbors, n_dims=3, ghost_mode=None, device=0):
        self._h = None
        desc = ContainerGridDesc()
        desc.n_dims = int(n_dims)
        for i in range(3):
            desc.dim[i] = int(dim[i])
            for s in range(2):
                desc.ghost_mode[i][s] = int(ghost_mode[i][s]) if ghost_mode is not None else GHOST_EDGE
        desc.capacity, desc.max_neighbors = int(capacity), int(max_neighbors)
        h = ctypes.c_void_p()
        check(lib().b200geo_containergrid_create(ctypes.byref(desc), int(device), ctypes.byref(h)))
        self._h = h
        self.dim, self.capacity, self.max_neighbors, self.device = tuple(dim), int(capacity), int(max_neighbors), device

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.b200geo_containergrid_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, arrays, origin=(0, 0, 0), dim=None, location=HOST, stream=None):
        dim = self.dim if dim is None else dim
        box = container_box(arrays)
        check(lib().b200geo_containergrid_load(self._h, _i3(origin), _i3(dim), ctypes.byref(box), location, stream))

    def save(self, arrays, origin=(0, 0, 0), dim=None, location=HOST, stream=None):
        dim = self.dim if dim is None else dim
        box = container_box(arrays)
        check(lib().b200geo_containergrid_save(self._h, _i3(origin), _i3(dim), ctypes.byref(box), location, stream))

    def set_edge(self, arrays):
        box = container_box(arrays)
        check(lib().b200geo_containergrid_set_edge(self._h, ctypes.byref(box)))

    def get_edge(self, arrays):
        box = container_box(arrays)
        check(lib().b200geo_containergrid_get_edge(self._h, ctypes.byref(box)))

    def step(self, kernel, n_steps=1, first_nano_step=0, params=None, stream=None):
        if kernel != KERNEL_CONTAINER:
            raise LogicError("a ContainerCell grid steps with KERNEL_CONTAINER")
        check(lib().b200geo_containergrid_step(self._h, first_nano_step, n_steps, stream))

    def stats_enable(self, on=True):
        pass

    def stats(self):
        out = (ctypes.c_uint64 * 6)()
        check(lib().b200geo_containergrid_stats(self._h, out))
        return {"cargo": int(out[0]), "links": int(out[1]), "resolutions": int(out[2]), "sweeps": int(out[3]),
                "kernel": int(out[4]), "link_table_bytes": int(out[5])}


def sync(stream=None):
    check(lib().b200geo_sync(stream))


def device_count():
    return lib().b200geo_device_count()


def set_tuning(key, value):
    check(lib().b200geo_set_tuning(key.encode(), int(value)))


def launch_count():
    return int(lib().b200geo_launch_count())
