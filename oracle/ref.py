"""ctypes view of oracle/_ref/libexags_ref.so: the REFERENCE's own gs code
(src/gs.upc, comm.upc, crystal.upc, sarray_transfer.upc, fail.upc, gs_local.c,
sort.c, sarray_sort.c compiled by build_ref.sh through the UPC shim), run as
np forked "UPC threads".  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libexags_ref.so")
ESZ = [8, 4, 4, 8, 8]
DT = [np.float64, np.float32, np.int32, np.int64, np.int64]


class Case(C.Structure):
    _fields_ = [("mode", C.c_int), ("vn", C.c_int), ("dom", C.c_int), ("op", C.c_int), ("transpose", C.c_int)]


def available() -> bool:
    return os.path.exists(LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_job.restype = C.c_int
        L.ref_job.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def run(ids_per_rank, cases, data, unique=0, method=1, want_maps=False, reps=0, arena_bytes=None):
    """cases: list of (mode, vn, dom, op, transpose) with the reference's enum
    values (Nek flavour: dom 0 double, 1 float, 2 int, 3 long, 4 long long).
    data[c][r]: ndarray of vn*n[r] elements (vec: tuples contiguous; many:
    [vn][n] contiguous), updated in place.  Returns (maps, times)."""
    np_ = len(ids_per_rank)
    ids = [np.ascontiguousarray(a, dtype=np.int64) for a in ids_per_rank]
    idtab = (C.c_void_p * np_)(*[a.ctypes.data for a in ids])
    ns = (C.c_uint * np_)(*[a.size for a in ids])
    cs = (Case * max(1, len(cases)))(*[Case(*c) for c in cases])
    dtab = (C.c_void_p * max(1, len(cases) * np_))(*[data[c][r].ctypes.data for c in range(len(cases)) for r in range(np_)])
    cap = 3 * max(a.size for a in ids) + 16
    maps = np.zeros((np_, cap), dtype=np.uint32) if want_maps else None
    mlen = np.zeros(np_ * 3, dtype=np.uint64)
    times = np.zeros(max(1, len(cases)))
    if arena_bytes is None:
        total = sum(a.size for a in ids)
        vmax = max([c[1] for c in cases] + [1])
        arena_bytes = (1 << 26) + total * (8 + 64 + 8 * vmax * max(1, len(cases))) * 2 + np_ * np_ * total * 2
    rc = lib().ref_job(np_, idtab, ns, unique, method, len(cases), cs, dtab,
                       maps.ctypes.data if want_maps else None, cap, mlen.ctypes.data, reps,
                       times.ctypes.data, int(arena_bytes))
    if rc != 0:
        raise RuntimeError("reference job failed")
    out_maps = None
    if want_maps:
        out_maps = []
        for r in range(np_):
            o, row = 0, []
            for w in range(3):
                L = int(mlen[r * 3 + w])
                row.append(maps[r, o:o + L].copy())
                o += L
            out_maps.append(row)
    return out_maps, times


def local(name: str, out, inp, vn: int, map_, dom: int, op: int) -> None:
    """The reference's own compiled gs_local.c entry point `name`
    (src/gs_local.h:23-41; Nek flavour symbols gslib_*), same calling convention
    as oracle.local.  Pure C, called in this process."""
    L = C.CDLL(LIB)

    def ptr(x):
        if isinstance(x, (list, tuple)):
            t = (C.c_void_p * len(x))(*[a.ctypes.data for a in x])
            return C.c_void_p(C.addressof(t)), t
        return C.c_void_p(x.ctypes.data), None
    po, k1 = ptr(out)
    f = getattr(L, "gslib_" + name)
    f.restype = None
    mp = C.c_void_p(map_.ctypes.data) if map_ is not None else None
    if name == "gs_gather_array":
        pi, k2 = ptr(inp)
        f(po, pi, C.c_uint(vn), C.c_int(dom), C.c_int(op))
    elif name == "gs_init_array":
        f(po, C.c_uint(vn), C.c_int(dom), C.c_int(op))
    elif "init" in name:
        f(po, C.c_uint(vn), mp, C.c_int(dom), C.c_int(op))
    elif "gather" in name:
        pi, k2 = ptr(inp)
        f(po, pi, C.c_uint(vn), mp, C.c_int(dom), C.c_int(op))
    else:
        pi, k2 = ptr(inp)
        f(po, pi, C.c_uint(vn), mp, C.c_int(dom))
