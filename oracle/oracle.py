"""ctypes view of oracle/libgs_oracle.so (the CPU restatement, gs_oracle.c) and
of oracle/_ref/ (the reference's own sources compiled by build_ref.sh).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product
(exaflow-exags-upc_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libgs_oracle.so")
REF_DIR = os.path.join(_HERE, "_ref")

D_DOUBLE, D_FLOAT, D_INT, D_LONG = 0, 1, 2, 3
O_ADD, O_MUL, O_MIN, O_MAX, O_BPR = 0, 1, 2, 3, 4
M_PLAIN, M_VEC, M_MANY = 0, 1, 2
METHOD_AUTO, METHOD_PAIRWISE, METHOD_CRYSTAL, METHOD_ALLREDUCE = 0, 1, 2, 3
NP_DOM = {np.dtype(np.float64): D_DOUBLE, np.dtype(np.float32): D_FLOAT,
          np.dtype(np.int32): D_INT, np.dtype(np.int64): D_LONG}


def build() -> None:
    r = subprocess.run(["make", "-C", _HERE, "libgs_oracle.so"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the oracle failed:\n" + r.stdout + r.stderr)
    if os.path.isdir("/root/reference/src"):
        r = subprocess.run(["make", "-C", _HERE, "ref"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + r.stdout + r.stderr)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.ogs_setup.restype = C.c_void_p
        L.ogs_setup.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ogs_free.argtypes = [C.c_void_p]
        L.ogs_total_shared.restype = C.c_ulonglong
        L.ogs_total_shared.argtypes = [C.c_void_p]
        L.ogs_map.restype = C.POINTER(C.c_uint)
        L.ogs_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.ogs_neighbours.restype = C.c_uint
        L.ogs_neighbours.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint]
        L.ogs_exec.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int]
        L.ogs_unique.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ogs_local.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


class Oracle:
    """All np ranks of one job, simulated in this process."""

    def __init__(self, ids_per_rank, unique: int = 0, method: int = METHOD_PAIRWISE):
        self.ids = [np.ascontiguousarray(a, dtype=np.int64) for a in ids_per_rank]
        self.np_ = len(self.ids)
        tab = (C.c_void_p * self.np_)(*[a.ctypes.data for a in self.ids])
        ns = (C.c_uint * self.np_)(*[a.size for a in self.ids])
        self.h = lib().ogs_setup(self.np_, tab, ns, unique, method)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ogs_free(self.h)
            self.h = None

    @property
    def total_shared(self) -> int:
        return int(lib().ogs_total_shared(self.h))

    def map(self, rank: int, which: int) -> np.ndarray:
        n = C.c_size_t()
        p = lib().ogs_map(self.h, rank, which, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.empty(0, np.uint32)

    def neighbours(self, rank: int, d: int):
        p = (C.c_uint * 256)()
        s = (C.c_uint * 256)()
        k = lib().ogs_neighbours(self.h, rank, d, p, s, 256)
        return [(int(p[i]), int(s[i])) for i in range(k)]

    def exec(self, us, mode: int, vn: int, dom: int, op: int, transpose: int) -> None:
        """us: one entry per rank; plain/vec: ndarray; many: list of vn ndarrays.  In place."""
        keep = []
        ptrs = []
        for u in us:
            if mode == M_MANY:
                t = (C.c_void_p * vn)(*[x.ctypes.data for x in u])
                keep.append(t)
                ptrs.append(C.addressof(t))
            else:
                ptrs.append(u.ctypes.data)
        tab = (C.c_void_p * self.np_)(*ptrs)
        lib().ogs_exec(self.h, tab, mode, vn, dom, op, transpose)


def unique(ids_per_rank) -> None:
    """gs_unique on every rank's ids, in place."""
    tab = (C.c_void_p * len(ids_per_rank))(*[a.ctypes.data for a in ids_per_rank])
    ns = (C.c_uint * len(ids_per_rank))(*[a.size for a in ids_per_rank])
    lib().ogs_unique(len(ids_per_rank), tab, ns)


# the public gs_local.h entry points (src/gs_local.h:23-41), in ogs_local's numbering
LOCAL_ENTRIES = ["gs_gather_array", "gs_init_array", "gs_gather", "gs_scatter", "gs_init",
                 "gs_gather_vec", "gs_scatter_vec", "gs_init_vec", "gs_gather_many", "gs_scatter_many",
                 "gs_init_many", "gs_gather_vec_to_many", "gs_scatter_many_to_vec", "gs_scatter_vec_to_many"]


def local(name: str, out, inp, vn: int, map_, dom: int, op: int) -> None:
    """One gs_local.h entry point on host arrays, in place on `out`.  out / inp:
    ndarray, or a list of vn ndarrays where the entry point takes an array of
    pointers; vn is the element count for the two *_array entries."""
    def ptr(x):
        if x is None:
            return None, None
        if isinstance(x, (list, tuple)):
            t = (C.c_void_p * len(x))(*[a.ctypes.data for a in x])
            return C.addressof(t), t
        return x.ctypes.data, None
    po, k1 = ptr(out)
    pi, k2 = ptr(inp)
    lib().ogs_local(LOCAL_ENTRIES.index(name), po, pi, vn, map_.ctypes.data if map_ is not None else None, dom, op)
