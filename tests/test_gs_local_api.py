"""The fourteen public gs_local.h entry points (src/gs_local.h:23-41, bodies
src/gs_local.c:187-331) as exported by libexags.so, on oracle-built sentinel
maps, every domain and op, host and device pointers -- against the oracle's
ogs_local, which tests/test_oracle_pin.py pins to the reference's compiled
gs_local.c.  Bit-exact."""
import copy
import ctypes as C
import itertools

import numpy as np
import pytest

import cases
import local_cases as LC

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _call(X, name, out, inp, vn, mp, dom, opi):
    L = X.lib()
    f = getattr(L, "exags_" + name)

    def ptr(x):
        if isinstance(x, list):
            t = (C.c_void_p * len(x))(*[X._addr(a) for a in x])
            return C.addressof(t), t
        return X._addr(x), None
    po, keep_o = ptr(out)
    mpp = mp.ctypes.data if mp is not None else None
    same = inp is out                       # the in-place form gs_aux uses
    if name == "gs_gather_array":
        pi, keep_i = ptr(inp)
        f(po, pi, vn, dom, opi)
    elif name == "gs_init_array":
        f(po, vn, dom, opi)
    elif "init" in name:
        f(po, vn, mpp, dom, opi)
    elif "gather" in name:
        pi, keep_i = (po, None) if same else ptr(inp)
        f(po, pi, vn, mpp, dom, opi)
    else:
        pi, keep_i = (po, None) if same else ptr(inp)
        f(po, pi, vn, mpp, dom)


def _to_dev(x):
    return [torch.from_numpy(a).cuda() for a in x] if isinstance(x, list) else torch.from_numpy(x).cuda()


def _to_host(x):
    return [a.cpu().numpy() for a in x] if isinstance(x, list) else x.cpu().numpy()


def _same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b)) if isinstance(a, list) else np.array_equal(a, b)


@pytest.fixture(scope="module")
def table(O):
    return LC.maps(O)


@pytest.mark.parametrize("where", ["host", "device"])
@pytest.mark.parametrize("dname", list(cases.DTYPES))
@pytest.mark.parametrize("name", list(LC.ENTRIES))
def test_local_entry_point(X, O, table, name, dname, where):
    dt = cases.DTYPES[dname]
    for op in cases.OPS_BPR:
        if not LC.ENTRIES[name][4] and op != "add":
            continue
        out, inp, vn, mp = LC.make(name, table, dt, op)
        w_out = copy.deepcopy(out)
        w_in = w_out if inp is out else copy.deepcopy(inp)
        dom, opi = O.NP_DOM[np.dtype(dt)], cases.OPS_BPR.index(op)
        O.local(name, w_out, w_in, vn, mp, dom, opi)
        g_out = copy.deepcopy(out)
        g_in = g_out if inp is out else copy.deepcopy(inp)
        if where == "device":
            d_out = _to_dev(g_out)
            d_in = d_out if inp is out else (_to_dev(g_in) if g_in is not None else None)
            _call(X, name, d_out, d_in, vn, mp, X.NP_DOM[np.dtype(dt)], opi)
            got, got_in = _to_host(d_out), (_to_host(d_in) if d_in is not None else None)
        else:
            _call(X, name, g_out, g_in, vn, mp, X.NP_DOM[np.dtype(dt)], opi)
            got, got_in = g_out, g_in
        assert _same(got, w_out), f"{name} {dname} {op} ({where} pointers)"
        if inp is not None and inp is not out:       # inputs are read-only
            assert _same(got_in, inp), f"{name} {dname} {op}: input modified"
