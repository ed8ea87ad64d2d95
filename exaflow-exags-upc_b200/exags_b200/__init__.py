"""exags_b200 -- Python view of libexags.so's C ABI (include/exags.h).

The product is the shared library: host C + hand-written sm_100a CUDA.  This
module only binds it with ctypes so that the parity tests and bench.py can
drive the same entry points a C or Fortran caller links against; it contains
no compute and no fallback.  Names and argument order mirror the reference
API (src/gs.h:129-138): gs_setup / gs / gs_vec / gs_many / gs_free / gs_unique.

`u` arguments may be numpy arrays (host pointers, staged through HBM by the
library), torch CUDA tensors, or raw device addresses (int).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("EXAGS_LIB_PATH") or os.path.join(_HERE, "libexags.so")   # override: variant builds for A/B runs

# gs_dom / gs_op / gs_method values (src/gs_defs.h:59-79, src/gs.h:127)
gs_double, gs_float, gs_int, gs_long, gs_long_long = 0, 1, 2, 3, 4
gs_add, gs_mul, gs_min, gs_max, gs_bpr = 0, 1, 2, 3, 4
gs_auto, gs_pairwise, gs_crystal_router, gs_all_reduce = 0, 1, 2, 3
mode_plain, mode_vec, mode_many = 0, 1, 2

DOM_NP = {gs_double: np.float64, gs_float: np.float32, gs_int: np.int32,
          gs_long: np.int64, gs_long_long: np.int64}
NP_DOM = {np.dtype(np.float64): gs_double, np.dtype(np.float32): gs_float,
          np.dtype(np.int32): gs_int, np.dtype(np.int64): gs_long}


def build(verbose: bool = False) -> str:
    """Compile libexags.so in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    r = subprocess.run(["make", "-C", _ROOT], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libexags.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stdout)
    return LIB_PATH


class comm(C.Structure):
    """struct comm (include/exags.h; first three members as src/comm.h:128-131)."""
    _fields_ = [("h", C.c_int), ("id", C.c_int), ("np", C.c_int),
                ("priv", C.c_void_p), ("ref_count", C.POINTER(C.c_int))]


comm_ptr = C.POINTER(comm)
_lib = None


def lib() -> C.CDLL:
    """Load the C-ABI library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C {_ROOT}` "
            "(or __graft_entry__.build()). There is no Python/CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u32, i32 = C.c_void_p, C.c_uint, C.c_int
    L.exags_comm_init.restype = None
    L.exags_comm_world.argtypes = [C.POINTER(comm_ptr)]
    L.exags_comm_free.argtypes = [C.POINTER(comm_ptr)]
    L.exags_comm_barrier.argtypes = [comm_ptr]
    L.exags_comm_allreduce.argtypes = [comm_ptr, i32, i32, vp, u32, vp]
    L.exags_comm_scan.argtypes = [vp, comm_ptr, i32, i32, vp, u32, vp]
    L.exags_comm_dot.argtypes = [comm_ptr, vp, vp, u32]
    L.exags_comm_dot.restype = C.c_double
    L.exags_comm_dot_async.argtypes = [comm_ptr, vp, vp, u32, vp, vp]
    L.exags_comm_dot_async.restype = None
    L.exags_comm_bcast.argtypes = [comm_ptr, vp, C.c_size_t, u32]
    L.gslib_gs_setup.argtypes = [vp, u32, comm_ptr, i32, i32, i32]
    L.gslib_gs_setup.restype = vp
    L.exags_gs_setup.argtypes = [vp, u32, comm_ptr, i32, i32, i32]
    L.exags_gs_setup.restype = vp
    for name in ("gslib_gs", "exags_gs"):
        getattr(L, name).argtypes = [vp, i32, i32, u32, vp, vp]
        getattr(L, name).restype = None
    for name in ("gslib_gs_vec", "exags_gs_vec", "gslib_gs_many", "exags_gs_many"):
        getattr(L, name).argtypes = [vp, u32, i32, i32, u32, vp, vp]
        getattr(L, name).restype = None
    L.exags_gs_async.argtypes = [vp, i32, u32, i32, i32, u32, vp, vp]
    L.exags_gs_async.restype = None
    L.exags_gs_weighted.argtypes = [vp, vp, i32, i32, vp]
    L.exags_gs_weighted.restype = None
    L.exags_gs_weighted_async.argtypes = [vp, vp, i32, i32, vp, vp]
    L.exags_gs_weighted_async.restype = None
    for name in ("exags_gs_weighted_vec", "exags_gs_weighted_many"):
        getattr(L, name).argtypes = [vp, u32, vp, i32, i32, vp]
        getattr(L, name).restype = None
        getattr(L, name + "_async").argtypes = [vp, u32, vp, i32, i32, vp, vp]
        getattr(L, name + "_async").restype = None
    L.gslib_gs_free.argtypes = [vp]
    L.exags_gs_free.argtypes = [vp]
    L.gslib_gs_unique.argtypes = [vp, u32, comm_ptr]
    L.exags_gs_unique.argtypes = [vp, u32, comm_ptr]
    L.exags_gs_info.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    L.exags_gs_dump_map.argtypes = [vp, i32, vp, C.c_size_t]
    L.exags_gs_dump_map.restype = C.c_size_t
    L.exags_kernel_launches.restype = C.c_ulonglong
    L.exags_exchange_transport.restype = C.c_char_p
    L.exags_set_device.argtypes = [i32]
    L.exags_get_device.restype = i32
    L.exags_device_malloc.argtypes = [C.c_size_t]
    L.exags_device_malloc.restype = vp
    L.exags_device_free.argtypes = [vp]
    L.exags_memcpy_h2d.argtypes = [vp, vp, C.c_size_t]
    L.exags_memcpy_d2h.argtypes = [vp, vp, C.c_size_t]
    L.exags_host_malloc_pinned.argtypes = [C.c_size_t]
    L.exags_host_malloc_pinned.restype = vp
    L.exags_host_free_pinned.argtypes = [vp]
    for name in ("exags_gs_gather", "exags_gs_gather_vec", "exags_gs_gather_many",
                 "exags_gs_gather_vec_to_many"):
        getattr(L, name).argtypes = [vp, vp, u32, vp, i32, i32]
        getattr(L, name).restype = None
    for name in ("exags_gs_scatter", "exags_gs_scatter_vec", "exags_gs_scatter_many",
                 "exags_gs_scatter_many_to_vec", "exags_gs_scatter_vec_to_many"):
        getattr(L, name).argtypes = [vp, vp, u32, vp, i32]
        getattr(L, name).restype = None
    for name in ("exags_gs_init", "exags_gs_init_vec", "exags_gs_init_many"):
        getattr(L, name).argtypes = [vp, u32, vp, i32, i32]
        getattr(L, name).restype = None
    L.exags_gs_gather_array.argtypes = [vp, vp, u32, i32, i32]
    L.exags_gs_init_array.argtypes = [vp, u32, i32, i32]
    _lib = L
    return L


def _addr(u) -> int:
    """Raw address of a numpy array, torch tensor, ctypes buffer or int."""
    if isinstance(u, int):
        return u
    if isinstance(u, np.ndarray):
        if not u.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return u.ctypes.data
    if hasattr(u, "data_ptr"):
        if hasattr(u, "is_contiguous") and not u.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return u.data_ptr()
    return C.addressof(u)


_world = None


def comm_world() -> comm_ptr:
    """comm_init + comm_world (src/comm.upc:35-49,223-253)."""
    global _world
    if _world is None:
        cp = comm_ptr()
        lib().exags_comm_world(C.byref(cp))
        _world = cp
    return _world


class GS:
    """A gs handle (struct gs_data *)."""

    def __init__(self, ptr: int, n: int, wide: bool):
        self.ptr, self.n, self.wide = ptr, n, wide

    def info(self) -> dict:
        a = (C.c_ulonglong * 16)()
        lib().exags_gs_info(self.ptr, a)
        keys = ["n", "groups", "members", "flagged_primaries", "neighbours", "send_slots",
                "recv_slots", "total_shared", "method", "sectors_f64", "boundary",
                "interior_groups", "interior_members", "flagged_singles", "handle_bytes",
                "sell_entries"]
        return {k: int(a[i]) for i, k in enumerate(keys)}

    def dump_map(self, which: int) -> np.ndarray:
        need = lib().exags_gs_dump_map(self.ptr, which, None, 0)
        out = np.empty(need, dtype=np.uint32)
        lib().exags_gs_dump_map(self.ptr, which, out.ctypes.data, need)
        return out


def gs_setup(ids, comm=None, unique: int = 0, method: int = gs_pairwise, verbose: int = 0) -> GS:
    """gs_setup(id, n, comm, unique, method, verbose) (src/gs.upc:1262-1269).
    int64 ids take the Nek flavour (gslib_, GLOBAL_LONG_LONG), int32 ids the
    default flavour (exags_)."""
    cp = comm if comm is not None else comm_world()
    if hasattr(ids, "data_ptr"):          # torch tensor: labels already in HBM (extension)
        if not ids.is_contiguous():
            raise ValueError("ids must be contiguous")
        size, itemsize = ids.numel(), ids.element_size()
    else:
        ids = np.ascontiguousarray(ids)
        if ids.dtype not in (np.int64, np.int32):
            raise TypeError("ids must be int32 or int64")
        size, itemsize = ids.size, ids.dtype.itemsize
    if itemsize == 8:
        p = lib().gslib_gs_setup(_addr(ids), size, cp, unique, method, verbose)
    elif itemsize == 4:
        p = lib().exags_gs_setup(_addr(ids), size, cp, unique, method, verbose)
    else:
        raise TypeError("ids must be int32 or int64")
    return GS(p, int(size), itemsize == 8)


def gs(u, dom: int, op: int, transpose: int, gsh: GS, buf=None) -> None:
    """gs(u, dom, op, transpose, gsh, buf) (src/gs.upc:1187-1191); in place."""
    lib().gslib_gs(_addr(u), dom, op, transpose, gsh.ptr, None)


def gs_vec(u, vn: int, dom: int, op: int, transpose: int, gsh: GS, buf=None) -> None:
    """gs_vec: u is T[n][vn] (src/gs.upc:1193-1197)."""
    lib().gslib_gs_vec(_addr(u), vn, dom, op, transpose, gsh.ptr, None)


def gs_many(us, vn: int, dom: int, op: int, transpose: int, gsh: GS, buf=None) -> None:
    """gs_many: us is a sequence of vn arrays T[n] (src/gs.upc:1199-1203)."""
    tab = (C.c_void_p * vn)(*[_addr(x) for x in us[:vn]])
    lib().gslib_gs_many(C.addressof(tab), vn, dom, op, transpose, gsh.ptr, None)


def gs_async(u, mode: int, vn: int, dom: int, op: int, transpose: int, gsh: GS, stream: int) -> None:
    """Stream-ordered gs on device-resident data (include/exags_cuda.h)."""
    if mode == mode_many:
        tab = (C.c_void_p * vn)(*[_addr(x) for x in u[:vn]])
        lib().exags_gs_async(C.addressof(tab), mode, vn, dom, op, transpose, gsh.ptr, stream)
    else:
        lib().exags_gs_async(_addr(u), mode, vn, dom, op, transpose, gsh.ptr, stream)


def gs_weighted(u, w, dom: int, op: int, gsh: GS, stream: int | None = None) -> None:
    """u <- w .* gs(u, op) in one pass (include/exags_cuda.h); device vectors."""
    if stream is None:
        lib().exags_gs_weighted(_addr(u), _addr(w), dom, op, gsh.ptr)
    else:
        lib().exags_gs_weighted_async(_addr(u), _addr(w), dom, op, gsh.ptr, stream)


def gs_weighted_vec(u, vn: int, w, dom: int, op: int, gsh: GS, stream: int | None = None) -> None:
    """u[i][k] <- w[i] * gs_vec(u, op)[i][k] in one pass (include/exags_cuda.h); device vectors."""
    if stream is None:
        lib().exags_gs_weighted_vec(_addr(u), vn, _addr(w), dom, op, gsh.ptr)
    else:
        lib().exags_gs_weighted_vec_async(_addr(u), vn, _addr(w), dom, op, gsh.ptr, stream)


def gs_weighted_many(us, vn: int, w, dom: int, op: int, gsh: GS, stream: int | None = None) -> None:
    """us[k][i] <- w[i] * gs_many(us, op)[k][i] in one pass (include/exags_cuda.h); device vectors."""
    tab = (C.c_void_p * vn)(*[_addr(x) for x in us[:vn]])
    if stream is None:
        lib().exags_gs_weighted_many(C.addressof(tab), vn, _addr(w), dom, op, gsh.ptr)
    else:
        lib().exags_gs_weighted_many_async(C.addressof(tab), vn, _addr(w), dom, op, gsh.ptr, stream)


def gs_free(gsh: GS) -> None:
    if gsh.ptr:
        lib().gslib_gs_free(gsh.ptr)
        gsh.ptr = None


def gs_unique(ids: np.ndarray, comm=None) -> None:
    """gs_unique(id, n, comm): negates non-owner ids in place (src/gs.upc:1281-1290)."""
    cp = comm if comm is not None else comm_world()
    if ids.dtype == np.int64:
        lib().gslib_gs_unique(ids.ctypes.data, ids.size, cp)
    elif ids.dtype == np.int32:
        lib().exags_gs_unique(ids.ctypes.data, ids.size, cp)
    else:
        raise TypeError("ids must be int32 or int64")


def exchange_transport() -> str:
    """'single' | 'push' | 'nccl' | 'ipc' (include/exags_cuda.h)."""
    return lib().exags_exchange_transport().decode()


def comm_dot(v, w, comm=None) -> float:
    """comm_dot(cp, v, w, n) (src/comm.upc:954): v, w host arrays or device tensors of doubles."""
    n = v.numel() if hasattr(v, "numel") else v.size
    return float(lib().exags_comm_dot(comm if comm is not None else comm_world(), _addr(v), _addr(w), n))


def comm_dot_async(v, w, out, stream: int, comm=None) -> None:
    """Stream-ordered comm_dot (include/exags_cuda.h): out (1 device double) = sum v*w over all ranks."""
    lib().exags_comm_dot_async(comm if comm is not None else comm_world(), _addr(v), _addr(w), v.numel(), _addr(out), stream)


def comm_allreduce(v, dom: int, op: int, comm=None) -> None:
    """comm_allreduce(cp, dom, op, v, vn, buf) in place (src/comm.upc:831)."""
    n = v.numel() if hasattr(v, "numel") else v.size
    lib().exags_comm_allreduce(comm if comm is not None else comm_world(), dom, op, _addr(v), n, None)


def comm_scan(v, dom: int, op: int, comm=None, out=None):
    """comm_scan(scan, cp, dom, op, v, vn, buf) (src/comm.upc:706): returns
    (exclusive prefix over lower ranks, total over all ranks).  v: host array or
    device tensor; out: optional device tensor / host array of 2*vn elements."""
    n = v.numel() if hasattr(v, "numel") else v.size
    if out is None:
        out = np.empty(2 * n, dtype=DOM_NP[dom])
    lib().exags_comm_scan(_addr(out), comm if comm is not None else comm_world(), dom, op, _addr(v), n, None)
    return out[:n], out[n:]


def comm_bcast(buf: np.ndarray, root: int, comm=None) -> None:
    """comm_bcast(cp, p, n, root) (src/comm.upc:427): bytes of a host array, in place."""
    lib().exags_comm_bcast(comm if comm is not None else comm_world(), buf.ctypes.data, buf.nbytes, root)


def kernel_launches() -> int:
    return int(lib().exags_kernel_launches())
