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


def test_install_layout_and_pkg_config(built, tmp_path):
    """`make install` lays out what the reference's does (src/Makefile.am:1-14,
    exags.pc.in): libexags, the headers under include/exags/, exags.pc.  A caller
    that uses the call patterns of the reference's in-tree users of gs
    (src/xxt.upc:619-631, src/amg.upc:94-109,200,215,266,403,468,543-545) builds
    and links from `pkg-config --cflags --libs exags` alone, in both flavours."""
    import shutil
    import subprocess
    if not shutil.which("pkg-config"):
        pytest.skip("pkg-config not installed")
    src = tmp_path / "intree_callers.c"
    src.write_text(r'''
#include <stddef.h>
#include "exags/c99.h"
#include "exags/name.h"
#include "exags/fail.h"
#include "exags/types.h"
#include "exags/comm.h"
#include "exags/mem.h"
#include "exags/gs_defs.h"
#include "exags/gs.h"
struct rid { sint p, i; };
int main(void) {
  comm_ptr world, comm; struct gs_data *gsh, *top; buffer buf = null_buffer;
  comm_init(); comm_world(&world); comm_dup(&comm, world);
  { ulong bid[4] = {3, 3, 5, 0}; sint v[4] = {6, 7, 1, 0};           /* xxt.upc:619-631 */
    gsh = gs_setup((const slong *)bid, 4, comm, 0, gs_pairwise, 0);
    gs(v, gs_sint, gs_bpr, 0, gsh, &buf);
    gs(v, gs_sint, gs_add, 0, gsh, &buf);
    gs_free(gsh); }
  { slong id[3] = {1, 1, 2}; double x[3] = {1, 2, 3}, avg[4] = {0}; struct rid rid[3] = {{0, 0}, {0, 1}, {0, 2}};
    int dump = 0, not_happy = 0; uint n = 3;
    comm_bcast(comm, &dump, sizeof(int), 0);                         /* amg.upc:543 */
    top = gs_setup((const slong *)id, n, comm, 1, gs_auto, 1);       /* amg.upc:545, 403 */
    gs(x, gs_double, gs_add, 1, top, 0);                             /* amg.upc:109,194 */
    gs(x, gs_double, gs_add, 0, top, 0);                             /* amg.upc:94,205 */
    avg[0] = comm_reduce_double(comm, gs_add, x, n);                 /* amg.upc:200 */
    comm_allreduce(comm, gs_double, gs_add, avg, 2, avg + 2);        /* amg.upc:215 */
    gs_vec(n ? &rid[0].p : 0, 2, gs_sint, gs_add, 0, top, &buf);     /* amg.upc:266 */
    if (comm_reduce_int(comm, gs_max, &not_happy, 1)) comm_barrier(comm);   /* amg.upc:468 */
    gs_free(top); }
  buffer_free(&buf); comm_free(&comm); comm_free(&world);
  return 0;
}''')
    for nek in ("", "1"):
        prefix = tmp_path / ("inst_nek" if nek else "inst")
        r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "exaflow-exags-upc_b200"), "install",
                            f"PREFIX={prefix}"] + (["NEK=1"] if nek else []), capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        for rel in ("lib/libexags.so", "lib/pkgconfig/exags.pc", "include/exags.h", "include/exags/gs.h",
                    "include/exags/comm.h", "include/exags/mem.h", "include/exags/types.h"):
            assert (prefix / rel).exists(), rel
        env = dict(os.environ, PKG_CONFIG_PATH=str(prefix / "lib" / "pkgconfig"))
        flags = subprocess.run(["pkg-config", "--cflags", "--libs", "exags"], capture_output=True, text=True, env=env)
        assert flags.returncode == 0, flags.stderr
        assert ("-DGLOBAL_LONG_LONG" in flags.stdout) == bool(nek)
        r = subprocess.run(["gcc", "-std=gnu99", "-Wall", "-Werror", str(src), "-o", str(tmp_path / "intree")]
                           + flags.stdout.split(), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_product_does_not_touch_the_oracle(built):
    """The oracle is test infrastructure: no source of the product names it, the
    built library neither links it nor carries a CPU reduction path by its name,
    and loading the product does not load it."""
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "exaflow-exags-upc_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".c", ".cu", ".h", ".py", ".map")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
    lib = os.path.join(pkg, "exags_b200", "libexags.so")
    needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
    assert "gs_oracle" not in needed and "exags_ref" not in needed
    code = ("import sys; sys.path.insert(0, %r); import exags_b200 as X; X.lib(); "
            "maps = open('/proc/self/maps').read(); "
            "assert 'libgs_oracle' not in maps and 'libexags_ref' not in maps; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % pkg)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
