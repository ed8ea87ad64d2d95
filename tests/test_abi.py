"""The C-ABI library loads and exports every function include/*.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in ("exags.h", "exags_cuda.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src):
            n = m.group(1)
            if n.startswith(("exags_", "gslib_", "fgslib_")):
                names.add(n)
    return sorted(names)


def test_library_exports_every_declared_symbol(X):
    L = ctypes.CDLL(X.LIB_PATH)
    names = declared_functions()
    assert len(names) > 60
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_fortran_and_data_symbols(X):
    L = ctypes.CDLL(X.LIB_PATH)
    for n in ("gs_setup_", "gs_op_", "gs_op_vec_", "gs_op_many_", "gs_op_fields_", "gs_free_",
              "gs_setup_pick_", "gs_op", "gs_op_vec", "gs_op_many", "gs_op_fields",
              "exags_glb_comm", "exags_comm_gbl_id", "exags_comm_gbl_np"):
        assert hasattr(L, n), n
    # communicator Fortran bindings (src/comm.upc:1031-1148), both flavours
    for n in ("init", "finalize", "world", "free", "np", "id", "type_int", "type_int8", "type_real",
              "type_dp", "tag_ub", "time", "barrier", "bcast", "allreduce_add", "allreduce_min",
              "allreduce_max", "allreduce_mul"):
        assert hasattr(L, f"comm_{n}_"), n
        assert hasattr(L, f"fgslib_comm_{n}_"), n


def test_fortran_comm_bindings_single_rank(X):
    """By-reference comm calls on one rank: identity, type tags, and an
    allreduce that must leave the vector unchanged (src/comm.upc:1062-1148)."""
    import numpy as np
    L = ctypes.CDLL(X.LIB_PATH)
    cp = ctypes.c_void_p()
    L.comm_init_()
    L.comm_world_(ctypes.byref(cp))
    npv, idv, ub = ctypes.c_int(-1), ctypes.c_int(-1), ctypes.c_int(-1)
    L.comm_np_(ctypes.byref(cp), ctypes.byref(npv))
    L.comm_id_(ctypes.byref(cp), ctypes.byref(idv))
    L.comm_tag_ub_(ctypes.byref(cp), ctypes.byref(ub))
    assert (npv.value, idv.value, ub.value) == (1, 0, 0)
    tags = []
    for n in ("int", "int8", "real", "dp"):
        t = ctypes.c_int(-1)
        getattr(L, f"comm_type_{n}_")(ctypes.byref(t))
        tags.append(t.value)
    assert len(set(tags)) == 4
    v = np.array([1.5, -2.0, 3.25]); buf = np.zeros(3)
    vn = ctypes.c_uint(3); dp = ctypes.c_int(tags[3])
    L.comm_allreduce_add_(ctypes.byref(cp), ctypes.byref(dp), v.ctypes.data_as(ctypes.c_void_p), ctypes.byref(vn),
                          buf.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(v, [1.5, -2.0, 3.25])
    L.comm_barrier_(ctypes.byref(cp))
    L.comm_free_(ctypes.byref(cp))
    assert not cp.value


def test_comm_reduce_single_rank(X):
    """comm_reduce__T folds a host array with op (src/comm.upc:961-984); n=0
    returns the identity of gs_defs.h:45-52."""
    import numpy as np
    L = ctypes.CDLL(X.LIB_PATH)
    cp = ctypes.c_void_p()
    L.exags_comm_init()
    L.exags_comm_world(ctypes.byref(cp))
    cases = (("double", ctypes.c_double, np.float64), ("float", ctypes.c_float, np.float32),
             ("int", ctypes.c_int, np.int32), ("long", ctypes.c_long, np.int64),
             ("long_long", ctypes.c_longlong, np.int64))
    for name, ct, dt in cases:
        f = getattr(L, f"exags_comm_reduce__{name}")
        f.restype = ct
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint]
        v = np.array([3, -2, 5, 4], dtype=dt)
        want = (v[0] + v[1] + v[2] + v[3], v[0] * v[1] * v[2] * v[3], v.min(), v.max())
        for op in range(4):
            assert f(cp, op, v.ctypes.data, 4) == want[op], (name, op)
        assert f(cp, X.gs_add, None, 0) == 0 and f(cp, X.gs_mul, None, 0) == 1
    L.exags_comm_free(ctypes.byref(cp))


def test_compat_headers_compile(tmp_path):
    """A caller written against the reference's header names compiles."""
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdlib.h>
#include "exags/c99.h"
#include "exags/name.h"
#include "exags/fail.h"
#include "exags/types.h"
#include "exags/comm.h"
#include "exags/mem.h"
#include "exags/gs_defs.h"
#include "exags/gs.h"
int caller(void) {
  comm_ptr cp; struct gs_data *gsh; buffer b = null_buffer;
  comm_init(); comm_world(&cp);
  slong *id = tmalloc(slong, 4); double v[4] = {1,1,1,1};
  id[0]=1; id[1]=1; id[2]=2; id[3]=0;
  gsh = gs_setup(id, 4, cp, 0, gs_pairwise, 0); free(id);
  gs(v, gs_double, gs_add, 0, gsh, &b);
  { sint one = 1; slong two = 2; (void)comm_reduce_sint(cp, gs_max, &one, 1); (void)comm_reduce_slong(cp, gs_add, &two, 1);
    (void)comm_reduce_double(cp, gs_add, v, 4); }
  gs_free(gsh); comm_free(&cp);
  return (int)cp->np;
}''')
    import subprocess
    for flags in ([], ["-DGLOBAL_LONG_LONG"]):
        r = subprocess.run(["gcc", "-std=gnu99", "-c", str(src), "-I", os.path.join(ROOT, "include"),
                            "-o", str(tmp_path / "caller.o")] + flags, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
