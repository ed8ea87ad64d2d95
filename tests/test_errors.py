"""Error behaviour of the C ABI (SURVEY.md 8b "Errors"): no return codes; a failure prints
`ERROR (proc NNNN, file:line): message` on STDOUT and exits with status 1, exactly the
reference's fail() contract (src/fail.upc:19-63).  Each case runs in its own process because
fail() ends it.  No GPU needed: these are the host-side checks, plus the two messages that
say there is no CPU execution path."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRE = f"""
import sys, ctypes as C
sys.path[:0] = [{os.path.join(ROOT, 'exaflow-exags-upc_b200')!r}]
import numpy as np
import exags_b200 as X
ids = np.array([1, 1, 2, 0, -2, 3], dtype=np.int64)
"""
ERR = re.compile(r"^ERROR \(proc \d{4}, [^:]+:\d+\): (.*)$", re.M)


def run(body, host_only=True):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.pop("EXAGS_HOST_SETUP_ONLY", None)
    if host_only:
        env["EXAGS_HOST_SETUP_ONLY"] = "1"
    else:
        env["CUDA_VISIBLE_DEVICES"] = ""          # a box with GPUs must behave like one without
    r = subprocess.run([sys.executable, "-c", PRE + body], capture_output=True, text=True, env=env, timeout=120)
    m = ERR.search(r.stdout)
    return r.returncode, (m.group(1) if m else None), r.stdout + r.stderr


CASES = [
    # (python body, substring of the message)
    ("g = X.gs_setup(ids)\nu = np.ones(6)\nX.gs(u, X.gs_double, X.gs_add, 0, g)", "no CPU execution path"),
    ("g = X.gs_setup(ids)\nX.lib().gslib_gs(0, 9, X.gs_add, 0, g.ptr, None)", "datatype 9 not in valid range"),
    ("g = X.gs_setup(ids)\nX.lib().gslib_gs(0, X.gs_double, 7, 0, g.ptr, None)", "op 7 not in valid range"),
    ("X.lib().gslib_gs(0, X.gs_double, X.gs_add, 0, None, None)", "null handle"),
    ("X.lib().gslib_gs_setup(ids.ctypes.data, 6, None, 0, 1, 0)", "null communicator"),
    ("X.lib().gslib_gs_setup(ids.ctypes.data, 6, X.comm_world(), 0, 17, 0)", "bad method 17"),
    # Fortran shims: 0-based handles, dom 1..3, op 1..4 (src/gs.upc:1346-1362)
    ("h, d, o, t = C.c_int(5), C.c_int(1), C.c_int(1), C.c_int(0)\n"
     "X.lib().fgslib_gs_op_(C.byref(h), None, C.byref(d), C.byref(o), C.byref(t))", "invalid handle"),
    ("h, d, o, t, n = C.c_int(-1), C.c_int(6), C.c_int(0), C.c_int(0), C.c_int(6)\n"
     "X.lib().fgslib_gs_setup_(C.byref(h), ids.ctypes.data, C.byref(n), C.byref(d), C.byref(d))\n"
     "d = C.c_int(4)\n"
     "X.lib().fgslib_gs_op_(C.byref(h), None, C.byref(d), C.byref(C.c_int(1)), C.byref(t))",
     "datatype 4 not in valid range 1-3"),
    ("h, d, o, t, n = C.c_int(-1), C.c_int(1), C.c_int(5), C.c_int(0), C.c_int(6)\n"
     "X.lib().fgslib_gs_setup_(C.byref(h), ids.ctypes.data, C.byref(n), C.byref(d), C.byref(d))\n"
     "X.lib().fgslib_gs_op_(C.byref(h), None, C.byref(d), C.byref(o), C.byref(t))", "op 5 not in valid range 1-4"),
    ("ct = C.c_int(99); v = np.ones(2); cp = X.comm_world()\n"
     "X.lib().exags_comm_allreduce_cdom(cp, ct, 0, v.ctypes.data, 2, v.ctypes.data)", "cannot identify cdom"),
]


@pytest.mark.parametrize("body,needle", CASES, ids=[c[1].split()[0] + str(i) for i, c in enumerate(CASES)])
def test_fail_contract(built, body, needle):
    code, msg, out = run(body)
    assert code == 1, out
    assert msg is not None and needle in msg, out


def test_setup_without_gpu_fails_loudly(built):
    """No device and no EXAGS_HOST_SETUP_ONLY: gs_setup itself refuses (there is no CPU fallback)."""
    code, msg, out = run("g = X.gs_setup(ids)", host_only=False)
    assert code == 1, out
    assert msg is not None and "no CUDA device available" in msg and "no CPU path" in msg, out


def test_missing_library_fails_loudly(built, tmp_path):
    """The Python binding has no fallback either: without libexags.so it raises."""
    body = ("import exags_b200 as Y\nY.LIB_PATH = %r\nY._lib = None\n"
            "try:\n    Y.lib()\nexcept RuntimeError as e:\n    print('RAISED', e); sys.exit(3)\n" % str(tmp_path / "nope.so"))
    env = dict(os.environ, EXAGS_HOST_SETUP_ONLY="1")
    r = subprocess.run([sys.executable, "-c", PRE + body], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 3 and "no Python/CPU fallback" in r.stdout, r.stdout + r.stderr
