"""Parity of the CUDA path against the CPU oracle, through the C ABI.
Bit-exact for every domain and op (one thread reduces a group left to right
in the reference's order, so even float/double add/mul agree exactly; the
north star allows 4 ulp there)."""
import ctypes as C
import itertools

import numpy as np
import pytest

import cases
from exags_b200 import semmesh

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).copy()).cuda()


def _run_both(X, O, g, o, dt, op, t, mode, vn, n, seed, host=False):
    opi = cases.OPS.index(op)
    dom, odom = X.NP_DOM[np.dtype(dt)], O.NP_DOM[np.dtype(dt)]
    if mode == 2:
        u = [cases.values(n, dt, op, seed + k) for k in range(vn)]
        want = [x.copy() for x in u]
        o.exec([want], O.M_MANY, vn, odom, opi, t)
        if host:
            got = [x.copy() for x in u]
            X.gs_many(got, vn, dom, opi, t, g)
        else:
            d = [_dev(x) for x in u]
            X.gs_many(d, vn, dom, opi, t, g)
            got = [x.cpu().numpy() for x in d]
        return np.stack(got), np.stack(want)
    u = cases.values(n * vn, dt, op, seed)
    want = u.copy()
    o.exec([want], mode, vn, odom, opi, t)
    if host:
        got = u.copy()
        (X.gs(got, dom, opi, t, g) if mode == 0 else X.gs_vec(got, vn, dom, opi, t, g))
    else:
        d = _dev(u)
        (X.gs(d, dom, opi, t, g) if mode == 0 else X.gs_vec(d, vn, dom, opi, t, g))
        got = d.cpu().numpy()
    return got, want


def test_known_answers_on_gpu(X):
    """outputs of the reference for a flagged label set."""
    g = X.gs_setup(cases.known_answer_ids())
    for t, want in ((0, [0, 12, 12, 12, 12, 14, 17, 14, 17, 17, 11, 12, 12, 12, 15, 15]),
                    (1, [1, 2, 14, 14, 14, 14, 26, 14, 9, 26, 11, 39, 13, 14, 31, 16])):
        d = _dev(np.arange(1, 17, dtype=np.float64))
        X.gs(d, X.gs_double, X.gs_add, t, g)
        assert d.cpu().numpy().tolist() == want
    h = np.arange(1, 17, dtype=np.float64)
    X.gs(h, X.gs_double, X.gs_max, 0, g)
    assert h.tolist() == [-np.finfo(np.float64).max, 5, 5, 5, 5, 8, 10, 8, 10, 10, 11, 12, 12, 12, 15, 15]
    X.gs_free(g)


def test_reference_fixture_np1(X):
    """tests/gs/gs_test.upc:138-144 at np=1."""
    g = X.gs_setup(cases.gs_test_fixture_ids(1, 0))
    for t, want in ((0, 3.0), (1, 4.0)):
        v = np.ones(5)
        X.gs(v, X.gs_double, X.gs_add, t, g)
        assert v[4] == want
    X.gs_free(g)


@pytest.mark.parametrize("dname,op", list(itertools.product(cases.DTYPES, cases.OPS)))
def test_all_domains_ops_flagged(X, O, dname, op):
    """every gs_dom x gs_op, transpose 0/1, plain/vec/many, on labels with
    zeros, flags, singletons and one 40-member group."""
    dt = cases.DTYPES[dname]
    ids = cases.random_ids(6000, 1500, 42, long_group=40)
    g, o = X.gs_setup(ids), O.Oracle([ids])
    for t, (mode, vn) in itertools.product((0, 1), ((0, 1), (1, 3), (2, 3), (1, 5), (2, 11))):
        got, want = _run_both(X, O, g, o, dt, op, t, mode, vn, ids.size, 7 * vn + t)
        assert np.array_equal(got, want), (dname, op, t, mode, vn)
    X.gs_free(g)


def test_host_pointers_staged(X, O):
    ids = cases.random_ids(3000, 700, 5)
    g, o = X.gs_setup(ids), O.Oracle([ids])
    for mode, vn in ((0, 1), (1, 3), (2, 3)):
        got, want = _run_both(X, O, g, o, np.float64, "add", 0, mode, vn, ids.size, 3, host=True)
        assert np.array_equal(got, want)
    X.gs_free(g)


@pytest.mark.parametrize("dname", list(cases.DTYPES))
def test_bpr_every_domain(X, O, dname):
    """gs_bpr (binary prefix, src/gs_defs.h:37-42; xxt.upc:623 issues it on sint):
    the reference instantiates it for every domain through a conversion to uint."""
    dt = cases.DTYPES[dname]
    ids = cases.random_ids(2000, 300, 9, p_flag=0.2)
    g, o = X.gs_setup(ids), O.Oracle([ids])
    for t, (mode, vn) in itertools.product((0, 1), ((0, 1), (1, 3), (2, 2))):
        if mode == 2:
            u = [cases.values(ids.size, dt, "bpr", 3 + k) for k in range(vn)]
            want = [x.copy() for x in u]
            o.exec([want], mode, vn, O.NP_DOM[np.dtype(dt)], O.O_BPR, t)
            d = [_dev(x) for x in u]
            X.gs_many(d, vn, X.NP_DOM[np.dtype(dt)], X.gs_bpr, t, g)
            assert all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(d, want))
            continue
        u = cases.values(ids.size * vn, dt, "bpr", 3)
        want = u.copy()
        o.exec([want], mode, vn, O.NP_DOM[np.dtype(dt)], O.O_BPR, t)
        d = _dev(u)
        (X.gs(d, X.NP_DOM[np.dtype(dt)], X.gs_bpr, t, g) if mode == 0
         else X.gs_vec(d, vn, X.NP_DOM[np.dtype(dt)], X.gs_bpr, t, g))
        assert np.array_equal(d.cpu().numpy(), want), f"{dname} t={t} mode={mode}"
    X.gs_free(g)


@pytest.mark.parametrize("ids", [np.zeros(0, np.int64), np.zeros(5, np.int64), np.array([-7], np.int64),
                                 np.arange(1, 50, dtype=np.int64)], ids=["empty", "allzero", "singleflag", "nogroups"])
def test_degenerate_inputs(X, O, ids):
    g, o = X.gs_setup(ids), O.Oracle([ids])
    for t in (0, 1):
        u = np.arange(1, ids.size + 1, dtype=np.float64)
        want = u.copy()
        o.exec([want], O.M_PLAIN, 1, O.D_DOUBLE, O.O_ADD, t)
        got = u.copy()
        if ids.size:
            X.gs(got, X.gs_double, X.gs_add, t, g)
        assert np.array_equal(got, want)
    X.gs_free(g)


def test_config1_small_sem_mesh(X, O):
    """BASELINE config 1: 2x2x2 hex N=3, add/min/max on double and int."""
    ids = semmesh.box_ids(2, 2, 2, 3)
    g, o = X.gs_setup(ids), O.Oracle([ids])
    v = _dev(np.ones(512))
    X.gs(v, X.gs_double, X.gs_add, 0, g)
    v = v.cpu().numpy()
    assert v.sum() == 1000 and v.max() == 8
    for dt, op in itertools.product((np.float64, np.int32), ("add", "min", "max")):
        got, want = _run_both(X, O, g, o, dt, op, 0, 0, 1, 512, 1)
        assert np.array_equal(got, want)
    iv = _dev(np.arange(512, dtype=np.int32))
    X.gs(iv, X.gs_int, X.gs_max, 0, g)
    assert int(iv[3]) == 64
    X.gs_free(g)


def test_unique_handle(X, O):
    ids = cases.random_ids(4000, 900, 13, p_flag=0.0)
    g, o = X.gs_setup(ids, unique=1), O.Oracle([ids], unique=1)
    for t in (0, 1):
        got, want = _run_both(X, O, g, o, np.float64, "add", t, 0, 1, ids.size, t)
        assert np.array_equal(got, want)
    X.gs_free(g)


def test_config3_vec_many_n7(X, O):
    """BASELINE config 3 shape: N=7 box, 3 components, float and double."""
    dims = (6, 5, 4, 7)
    ids = semmesh.box_ids(*dims)
    g, o = X.gs_setup(ids), O.Oracle([ids])
    for dt, (mode, vn) in itertools.product((np.float32, np.float64), ((1, 3), (2, 3))):
        got, want = _run_both(X, O, g, o, dt, "add", 0, mode, vn, ids.size, 17)
        assert np.array_equal(got, want)
    X.gs_free(g)


def test_medium_mesh_vs_oracle(X, O):
    """16x16x8 N=7 (1.05 M DOF): the largest size the oracle sets up in seconds."""
    dims = (16, 16, 8, 7)
    ids = semmesh.box_ids(*dims)
    dof = semmesh.slab_dof_numbers(*dims)
    g, o = X.gs_setup(ids), O.Oracle([ids])
    for dt, op in ((np.float64, "add"), (np.float32, "mul"), (np.int64, "min"), (np.int32, "max")):
        u = semmesh.field_values(dof, dt, op)
        want = u.copy()
        o.exec([want], O.M_PLAIN, 1, O.NP_DOM[np.dtype(dt)], cases.OPS.index(op), 0)
        d = _dev(u)
        X.gs(d, X.NP_DOM[np.dtype(dt)], cases.OPS.index(op), 0, g)
        assert np.array_equal(d.cpu().numpy(), want)
    X.gs_free(g)


def test_full_size_properties(X):
    """BASELINE config 2 at full size (M7, 100 663 296 DOF): size-independent
    properties.  gs(add) of ones is the multiplicity m; sum(1/m) counts the
    unique nodes (exact: m is a power of two); gs(max) is idempotent; gs(add)
    is linear under exact scalings; the field is symmetric afterwards."""
    dims = (64, 64, 48, 7)
    c = semmesh.box_counts(*dims)
    ids = semmesh.box_ids(*dims)
    g = X.gs_setup(ids)
    info = g.info()
    assert (info["n"], info["groups"], info["members"]) == (c["n"], c["groups"], c["members"])
    # at N=7 every 32 B sector holds a shared node except along the outer faces of the box
    assert 0.99 * (c["n"] // 4) < info["sectors_f64"] <= c["n"] // 4
    idt = torch.from_numpy(ids).cuda()
    del ids
    m = torch.ones(c["n"], dtype=torch.float64, device="cuda")
    X.gs(m, X.gs_double, X.gs_add, 0, g)
    assert set(torch.unique(m).tolist()) == {1.0, 2.0, 4.0, 8.0}
    assert float((1.0 / m).sum()) == float(c["unique"])
    # every copy of a node holds the same value after gs: compare with a
    # scatter-reduce by global id done by torch on the same GPU
    x = torch.rand(c["n"], dtype=torch.float64, device="cuda") + 0.5
    y = x.clone()
    X.gs(y, X.gs_double, X.gs_max, 0, g)
    ref = torch.full((c["unique"] + 1,), -1.0, dtype=torch.float64, device="cuda")
    ref.scatter_reduce_(0, idt, x, reduce="amax")
    assert torch.equal(y, ref[idt])
    y2 = y.clone()
    X.gs(y2, X.gs_double, X.gs_max, 0, g)
    assert torch.equal(y, y2)                          # idempotent
    a = x.clone(); X.gs(a, X.gs_double, X.gs_add, 0, g)
    b = (4.0 * x); X.gs(b, X.gs_double, X.gs_add, 0, g)
    assert torch.equal(b, 4.0 * a)                     # linear under exact scaling
    iv = torch.randint(-1000, 1001, (c["n"],), dtype=torch.int64, device="cuda")
    want = torch.zeros(c["unique"] + 1, dtype=torch.int64, device="cuda").scatter_add_(0, idt, iv)[idt]
    X.gs(iv, X.gs_long, X.gs_add, 0, g)
    assert torch.equal(iv, want)                       # integer add is exact
    X.gs_free(g)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32, torch.int32])
def test_config3_full_size(X, dt):
    """BASELINE config 3 at full size (M7, 3 components): gs_vec and gs_many
    give, component by component, exactly what three plain gs calls give (same
    association order) -- a size-independent property for sizes the CPU oracle
    cannot reach (src/gs_local.c:114-158,256-288 against :60-97)."""
    dom = {torch.float64: X.gs_double, torch.float32: X.gs_float, torch.int32: X.gs_int}[dt]
    ids = semmesh.box_ids(64, 64, 48, 7)
    n = ids.size
    g = X.gs_setup(ids)
    del ids
    gen = torch.Generator(device="cuda").manual_seed(7)
    if dt.is_floating_point:
        comps = [torch.rand(n, dtype=dt, device="cuda", generator=gen) + 0.5 for _ in range(3)]
    else:
        comps = [torch.randint(-1000, 1001, (n,), dtype=dt, device="cuda", generator=gen) for _ in range(3)]
    for op in (X.gs_add, X.gs_max):
        want = [c.clone() for c in comps]
        for c in want:
            X.gs(c, dom, op, 0, g)
        many = [c.clone() for c in comps]
        X.gs_many(many, 3, dom, op, 0, g)
        assert all(torch.equal(a, b) for a, b in zip(many, want)), "gs_many != 3 x gs"
        del many
        vec = torch.stack(comps, dim=1).contiguous().view(-1)
        X.gs_vec(vec, 3, dom, op, 0, g)
        assert torch.equal(vec.view(n, 3), torch.stack(want, dim=1)), "gs_vec != 3 x gs"
        del vec
        if dt.is_floating_point and op == X.gs_add:
            w = torch.rand(n, dtype=dt, device="cuda", generator=gen) * 0.75 + 0.25
            m = torch.ones(n, dtype=dt, device="cuda")
            X.gs(m, dom, X.gs_add, 0, g)
            fused = comps[0].clone()
            X.gs_weighted(fused, w, dom, X.gs_add, g)
            assert torch.equal(fused, torch.where(m > 1, want[0] * w, want[0])), "gs_weighted != gs then *w"
            del w, m, fused
        del want
    X.gs_free(g)


def test_map_too_large_for_sliced_layout_falls_back_to_csr(X, O, monkeypatch):
    """Beyond the sliced layout's size limits (2^30 member indices) gs_setup keeps
    the map in CSR form and the thread-per-group kernel serves it; the limit is
    lowered here so that a small map takes that path."""
    ids = cases.random_ids(4000, 700, 12)
    monkeypatch.setenv("EXAGS_SELL_MAX_MEMBERS", "100")
    g, o = X.gs_setup(ids), O.Oracle([ids])
    monkeypatch.delenv("EXAGS_SELL_MAX_MEMBERS")
    assert g.info()["sell_entries"] == 0 and g.info()["interior_groups"] > 0
    for (dname, dt), op, t, (mode, vn) in itertools.product(cases.DTYPES.items(), ("add", "min"), (0, 1),
                                                             ((0, 1), (1, 3), (2, 3))):
        got, want = _run_both(X, O, g, o, dt, op, t, mode, vn, ids.size, 5)
        assert np.array_equal(got, want), (dname, op, t, mode)
    host = cases.values(ids.size, np.float64, "add", 9)
    want = host.copy()
    o.exec([want], O.M_PLAIN, 1, O.D_DOUBLE, O.O_ADD, 0)
    X.gs(host, X.gs_double, X.gs_add, 0, g)
    assert np.array_equal(host, want)
    X.gs_free(g)


def test_fortran_bindings(X):
    """fgs_setup / gs_op / gs_op_vec / gs_op_many / gs_op_fields / gs_free by reference
    (src/gs.upc:1309-1418), Nek flavour: dom 1..3, op 1..4, 0-based handles."""
    L = X.lib()
    ids = cases.random_ids(500, 100, 3, p_flag=0.0)
    n = C.c_int(ids.size); h = C.c_int(-1); zero = C.c_int(0)
    L.fgslib_gs_setup_(C.byref(h), ids.ctypes.data_as(C.c_void_p), C.byref(n), C.byref(zero), C.byref(zero))
    assert h.value >= 0
    g = X.gs_setup(ids)
    i = lambda v: C.byref(C.c_int(v))
    u = cases.values(ids.size, np.float64, "add", 1); want = u.copy()
    X.gs(want, X.gs_double, X.gs_add, 0, g)
    L.fgslib_gs_op_(C.byref(h), u.ctypes.data_as(C.c_void_p), i(1), i(1), i(0))
    assert np.array_equal(u, want)
    u = cases.values(ids.size, np.int64, "max", 2); want = u.copy()
    X.gs(want, X.gs_long, X.gs_max, 0, g)
    L.fgslib_gs_op_(C.byref(h), u.ctypes.data_as(C.c_void_p), i(3), i(4), i(0))
    assert np.array_equal(u, want)
    u = cases.values(ids.size * 2, np.float64, "add", 3); want = u.copy()
    X.gs_vec(want, 2, X.gs_double, X.gs_add, 1, g)
    L.fgslib_gs_op_vec_(C.byref(h), u.ctypes.data_as(C.c_void_p), i(2), i(1), i(1), i(1))
    assert np.array_equal(u, want)
    f = cases.values(ids.size * 3, np.float64, "add", 4); want = f.copy()
    X.gs_many([want[k * ids.size:(k + 1) * ids.size] for k in range(3)], 3, X.gs_double, X.gs_add, 0, g)
    L.fgslib_gs_op_fields_(C.byref(h), f.ctypes.data_as(C.c_void_p), i(ids.size), i(3), i(1), i(1), i(0))
    assert np.array_equal(f, want)
    L.fgslib_gs_free_(C.byref(h))
    X.gs_free(g)


def test_gs_local_api(X):
    """Public gs_local.h kernels on sentinel maps (src/gs_local.h:23-41) against
    a plain Python statement of src/gs_local.c:60-97."""
    NIL = 0xFFFFFFFF
    mp = np.array([2, 3, 4, 1, NIL, 5, 7, NIL, 6, 9, 8, NIL, NIL], dtype=np.uint32)
    u = np.arange(1, 13, dtype=np.float64)
    want = u.copy()
    for grp in ([2, 3, 4, 1], [5, 7], [6, 9, 8]):
        for j in grp[1:]:
            want[grp[0]] += u[j]
    got = u.copy()
    X.lib().exags_gs_gather(got.ctypes.data, got.ctypes.data, 1, mp.ctypes.data, X.gs_double, X.gs_add)
    assert np.array_equal(got, want)
    want2 = want.copy()
    for grp in ([2, 3, 4, 1], [5, 7], [6, 9, 8]):
        for j in grp[1:]:
            want2[j] = want2[grp[0]]
    X.lib().exags_gs_scatter(got.ctypes.data, got.ctypes.data, 1, mp.ctypes.data, X.gs_double)
    assert np.array_equal(got, want2)
    lst = np.array([0, 5, NIL], dtype=np.uint32)
    X.lib().exags_gs_init(got.ctypes.data, 1, lst.ctypes.data, X.gs_double, X.gs_min)
    assert got[0] == np.finfo(np.float64).max and got[5] == np.finfo(np.float64).max
    a, b = np.arange(6, dtype=np.int32), np.arange(6, dtype=np.int32)[::-1].copy()
    X.lib().exags_gs_gather_array(a.ctypes.data, b.ctypes.data, 6, X.gs_int, X.gs_max)
    assert a.tolist() == [5, 4, 3, 3, 4, 5]
    X.lib().exags_gs_init_array(a.ctypes.data, 6, X.gs_int, X.gs_mul)
    assert a.tolist() == [1] * 6
    # many -> vec pack and vec -> many unpack (pairwise buffer layout [slot][vn])
    pm = np.array([3, 0, NIL, 5, 1, NIL, NIL], dtype=np.uint32)     # u[3]->buf[0], u[5]->buf[1]
    us = [np.arange(10, 16, dtype=np.float64) + 100 * k for k in range(2)]
    tab = (C.c_void_p * 2)(*[x.ctypes.data for x in us])
    buf = np.zeros(4)
    X.lib().exags_gs_scatter_many_to_vec(buf.ctypes.data, C.addressof(tab), 2, pm.ctypes.data, X.gs_double)
    assert buf.tolist() == [13, 113, 15, 115]
    X.lib().exags_gs_gather_vec_to_many(C.addressof(tab), buf.ctypes.data, 2, pm.ctypes.data, X.gs_double, X.gs_add)
    assert us[0][3] == 26 and us[1][3] == 226 and us[0][5] == 30 and us[1][5] == 230


def test_kernels_really_launch(X):
    ids = cases.random_ids(1000, 200, 1)
    g = X.gs_setup(ids)
    before = X.kernel_launches()
    X.gs(_dev(np.ones(ids.size)), X.gs_double, X.gs_add, 0, g)
    assert X.kernel_launches() > before
    X.gs_free(g)


import glob as _glob
import os as _os

_GOLD1 = sorted(p for p in _glob.glob(_os.path.join(_os.path.dirname(__file__), "golden", "*.npz"))
                if "_np1" in p or "config1" in p)


@pytest.mark.parametrize("path", _GOLD1, ids=[_os.path.basename(p) for p in _GOLD1])
def test_golden_from_reference_on_gpu(X, path):
    """The CUDA path against fixtures produced by the reference's own compiled
    sources (oracle/make_golden.py): bit-exact, every case."""
    z = np.load(path)
    ids = z["id0"]
    g = X.gs_setup(ids, unique=int(z["unique"]))
    for w in (0, 1, 2):
        assert np.array_equal(g.dump_map(w), z[f"map{w}_0"])
    k = 0
    while f"case{k}_meta" in z:
        mode, vn, dom, op, t = [int(x) for x in z[f"case{k}_meta"]]
        a = z[f"case{k}_in0"].copy()
        if mode == 2:
            d = [_dev(a[i]) for i in range(vn)]
            X.gs_many(d, vn, dom, op, t, g)
            got = np.stack([x.cpu().numpy() for x in d])
        else:
            d = _dev(a)
            (X.gs(d, dom, op, t, g) if mode == 0 else X.gs_vec(d, vn, dom, op, t, g))
            got = d.cpu().numpy()
        assert np.array_equal(got, z[f"case{k}_out0"]), (path, k)
        k += 1
    X.gs_free(g)


def test_pipelined_host_path(X):
    """Host pointers on a field large enough to take the chunked
    H2D / wave / D2H pipeline (>= 48 MB of doubles): must equal the
    device-resident result exactly, for plain, vec and many."""
    dims = (32, 32, 16, 7)
    ids = semmesh.box_ids(*dims)
    n = ids.size
    g = X.gs_setup(ids)
    rng = np.random.default_rng(3)
    u = rng.uniform(0.5, 1.5, n)
    d = _dev(u); X.gs(d, X.gs_double, X.gs_add, 0, g)
    h = u.copy(); X.gs(h, X.gs_double, X.gs_add, 0, g)
    assert np.array_equal(h, d.cpu().numpy())
    v = rng.uniform(0.5, 1.5, 3 * n).astype(np.float32)
    d = _dev(v); X.gs_vec(d, 3, X.gs_float, X.gs_add, 1, g)
    h = v.copy(); X.gs_vec(h, 3, X.gs_float, X.gs_add, 1, g)
    assert np.array_equal(h, d.cpu().numpy())
    m = [rng.integers(-1000, 1000, n) for _ in range(3)]
    dm = [_dev(a) for a in m]; X.gs_many(dm, 3, X.gs_long, X.gs_max, 0, g)
    hm = [a.copy() for a in m]; X.gs_many(hm, 3, X.gs_long, X.gs_max, 0, g)
    for a, b in zip(hm, dm):
        assert np.array_equal(a, b.cpu().numpy())
    X.gs_free(g)


@pytest.mark.parametrize("dt", [np.float64, np.float32, np.int32, np.int64])
def test_vec3_aligned_and_unaligned(X, O, dt):
    """gs_vec with 3-tuples takes the pair+single load kernel when the field is
    aligned to two elements and the generic one when it is not (a view that
    starts one element into an allocation); both must equal the oracle."""
    ids = semmesh.box_ids(5, 4, 3, 4)
    n = ids.size
    g = X.gs_setup(ids)
    o = O.Oracle([ids])
    dom, odom = X.NP_DOM[np.dtype(dt)], O.NP_DOM[np.dtype(dt)]
    for op in ("add", "max"):
        opi = cases.OPS.index(op)
        u = cases.values(3 * n, dt, op, 77)
        want = u.copy()
        o.exec([want], O.M_VEC, 3, odom, opi, 0)
        d = _dev(u)
        X.gs_vec(d, 3, dom, opi, 0, g)
        assert np.array_equal(d.cpu().numpy(), want), ("aligned", op)
        big = torch.zeros(3 * n + 1, dtype=d.dtype, device="cuda")
        view = big[1:]
        view.copy_(torch.from_numpy(u))
        assert view.data_ptr() % (2 * view.element_size()) != 0
        X.gs_vec(view.data_ptr(), 3, dom, opi, 0, g)
        assert np.array_equal(view.cpu().numpy(), want), ("unaligned", op)
    X.gs_free(g)


def test_flagged_large_sliced(X, O):
    """A map with flagged members large enough to fill many sliced-ELL windows
    (flag mask path of the hot kernel), plain / vec / many, both transposes."""
    ids = semmesh.box_ids(8, 8, 6, 5)
    ids = ids.copy()
    rng = np.random.default_rng(11)
    flip = rng.random(ids.size) < 0.3
    ids[flip] = -ids[flip]
    ids[rng.random(ids.size) < 0.02] = 0
    g = X.gs_setup(ids)
    o = O.Oracle([ids])
    for dt, op, t, mode, vn in ((np.float64, "add", 0, 0, 1), (np.float64, "add", 1, 0, 1),
                                (np.float32, "mul", 1, 1, 3), (np.int32, "min", 0, 2, 3),
                                (np.int64, "max", 1, 2, 2), (np.float64, "max", 0, 1, 2)):
        got, want = _run_both(X, O, g, o, dt, op, t, mode, vn, ids.size, 5)
        assert np.array_equal(got, want), (dt, op, t, mode)
    X.gs_free(g)


def test_c_caller_links_and_runs(X, tmp_path):
    """A C program written against the reference's headers and spellings, linked
    with -lexags and run: the drop-in boundary exercised without Python in between."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(X.LIB_PATH)
    exe = tmp_path / "caller_run"
    for flags in ([], ["-DGLOBAL_LONG_LONG"]):
        r = subprocess.run(["gcc", "-std=gnu99", "-O1", os.path.join(root, "tests", "c", "caller_run.c"),
                            "-I", os.path.join(root, "include"), "-L", libdir, "-lexags",
                            f"-Wl,-rpath,{libdir}", "-o", str(exe)] + flags, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "EXAGS_HOST_SETUP_ONLY")}
        r = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=120)
        assert r.returncode == 0 and "caller_run ok" in r.stdout, r.stdout + r.stderr


def test_comm_dot_device_vectors(X):
    """comm_dot on device vectors: exact when every partial sum is an integer
    below 2^53 (any association gives the same bits), within rounding otherwise;
    odd lengths and the empty vector included."""
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 4096, 1_000_003):
        v = rng.integers(-8, 9, n).astype(np.float64)
        w = rng.integers(-8, 9, n).astype(np.float64)
        want = float(np.dot(v, w))
        assert X.comm_dot(v, w) == want                     # host vectors (reference path)
        if n:
            assert X.comm_dot(_dev(v), _dev(w)) == want     # device vectors
    v, w = rng.uniform(-1, 1, 3_000_001), rng.uniform(-1, 1, 3_000_001)
    got, want = X.comm_dot(_dev(v), _dev(w)), float(np.dot(v, w))
    scale = float(np.dot(np.abs(v), np.abs(w)))
    assert abs(got - want) <= 64 * np.finfo(np.float64).eps * scale


def _shared_mask(ids):
    """DOFs whose label is shared with another DOF (members of a group)."""
    labs = np.abs(ids)
    uniq, cnt = np.unique(labs[labs != 0], return_counts=True)
    return np.isin(labs, uniq[cnt >= 2])


@pytest.mark.parametrize("kind", ["sem", "random_flagged", "long_groups"])
def test_weighted_gs(X, O, kind):
    """exags_gs_weighted (include/exags_cuda.h): Nek's dssum followed by the
    multiplication with an inverse-multiplicity weight, fused into the one pass.
    Equal, bit for bit, to gs followed by u[i] *= w[i] on every member of a group
    (one rounding per product either way); other DOFs untouched.  Sliced kernel
    without and with the flag mask, and the CSR kernel of long groups."""
    if kind == "sem":
        ids = semmesh.box_ids(6, 5, 4, 5)
    elif kind == "random_flagged":
        ids = cases.random_ids(20000, 6000, 21)
    else:
        ids = cases.random_ids(9000, 400, 22, p_flag=0.0, long_group=60)
    n = ids.size
    shared = _shared_mask(ids)
    g, o = X.gs_setup(ids), O.Oracle([ids])
    for dt, op in itertools.product((np.float64, np.float32), cases.OPS):
        opi = cases.OPS.index(op)
        u = cases.values(n, dt, op, 31)
        w = np.random.default_rng(8).uniform(0.25, 1.0, n).astype(dt)
        want = u.copy()
        o.exec([want], O.M_PLAIN, 1, O.NP_DOM[np.dtype(dt)], opi, 0)
        with np.errstate(over="ignore"):
            want = np.where(shared, want * w, want).astype(dt)
        d, dw = _dev(u), _dev(w)
        X.gs_weighted(d, dw, X.NP_DOM[np.dtype(dt)], opi, g)
        assert np.array_equal(d.cpu().numpy(), want), (kind, dt, op)
        # stream-ordered form on torch's current stream
        d2 = _dev(u)
        X.gs_weighted(d2, dw, X.NP_DOM[np.dtype(dt)], opi, g, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d2.cpu().numpy(), want), (kind, dt, op, "async")
    X.gs_free(g)


@pytest.mark.parametrize("kind", ["sem", "random_flagged", "long_groups"])
def test_weighted_gs_vec_many(X, kind):
    """exags_gs_weighted_vec / _many (one weight per node for all components, a
    velocity field): bit for bit what vn weighted plain calls give -- and those are
    checked against the oracle in test_weighted_gs."""
    ids = {"sem": semmesh.box_ids(5, 4, 3, 7), "random_flagged": cases.random_ids(6000, 900, 5),
           "long_groups": cases.random_ids(4000, 60, 6, p_flag=0.0, long_group=50)}[kind]
    n = ids.size
    g = X.gs_setup(ids)
    for dt, vn, op in itertools.product((np.float64, np.float32), (2, 3, 5), ("add", "max")):
        dom, opi = X.NP_DOM[np.dtype(dt)], cases.OPS.index(op)
        comps = [cases.values(n, dt, op, 30 + k) for k in range(vn)]
        w = _dev(np.random.default_rng(8).uniform(0.25, 1.0, n).astype(dt))
        want = []
        for c in comps:
            d = _dev(c)
            X.gs_weighted(d, w, dom, opi, g)
            want.append(d)
        many = [_dev(c) for c in comps]
        X.gs_weighted_many(many, vn, w, dom, opi, g)
        assert all(torch.equal(a, b) for a, b in zip(many, want)), (kind, dt, vn, op, "many")
        vec = _dev(np.stack(comps, axis=1).reshape(-1))
        X.gs_weighted_vec(vec, vn, w, dom, opi, g)
        assert torch.equal(vec.view(n, vn), torch.stack(want, dim=1)), (kind, dt, vn, op, "vec")
    X.gs_free(g)


def test_weighted_gs_is_dsavg(X):
    """With w = 1 / multiplicity the weighted sum is the average over the copies
    of a node: a field that is already continuous is reproduced exactly."""
    dims = (8, 8, 6, 7)
    ids = semmesh.box_ids(*dims)
    g = X.gs_setup(ids)
    m = torch.ones(ids.size, dtype=torch.float64, device="cuda")
    X.gs(m, X.gs_double, X.gs_add, 0, g)
    w = 1.0 / m                                            # multiplicities are 1, 2, 4, 8: exact
    cont = torch.from_numpy(((ids * 2654435761) % 1024).astype(np.float64)).cuda()   # a function of the node
    u = cont.clone()
    X.gs_weighted(u, w, X.gs_double, X.gs_add, g)
    assert torch.equal(u, cont)
    X.gs_free(g)


def test_comm_dot_async(X):
    """Stream-ordered comm_dot (include/exags_cuda.h): same bits as the blocking
    call, result left on the device, capturable in a CUDA graph together with gs."""
    n = 1_000_003
    rng = np.random.default_rng(2)
    a, b = _dev(rng.uniform(-1, 1, n + 1)[:n]), _dev(rng.uniform(-1, 1, n + 1)[:n])
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    X.comm_dot_async(a, b, out, s)
    torch.cuda.synchronize()
    assert float(out.item()) == X.comm_dot(a, b)
    assert abs(float(out.item()) - float(torch.dot(a, b).item())) <= 1e-9 * n ** 0.5
    ints = _dev(rng.integers(-8, 9, n).astype(np.float64))
    X.comm_dot_async(ints, ints, out, s)
    torch.cuda.synchronize()
    assert float(out.item()) == float((ints * ints).sum().item())      # exact in any order
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        X.comm_dot_async(a, b, out, torch.cuda.current_stream().cuda_stream)
    out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert float(out.item()) == X.comm_dot(a, b)


def test_comm_scan_device_vectors(X):
    """comm_scan with device input and output (one rank: empty prefix, own total)."""
    v = torch.tensor([5, -2, 9], dtype=torch.int64, device="cuda")
    out = torch.zeros(6, dtype=torch.int64, device="cuda")
    X.comm_scan(v, X.gs_long, X.gs_add, out=out)
    assert out.cpu().tolist() == [0, 0, 0, 5, -2, 9]
    excl, tot = X.comm_scan(np.array([2.5, 4.0]), X.gs_double, X.gs_max)
    assert excl.tolist() == [-np.finfo(np.float64).max] * 2 and tot.tolist() == [2.5, 4.0]


def test_gs_async_under_cuda_graph(X, O):
    """exags_gs_async with one rank makes no host synchronisation and no
    allocation: a launch-bound inner loop of small gs calls can be captured in a
    CUDA graph and replayed (the captured kernels read the same device map)."""
    ids = semmesh.box_ids(4, 4, 3, 5)
    g, o = X.gs_setup(ids), O.Oracle([ids])
    u0 = cases.values(ids.size, np.float64, "add", 4)
    want = u0.copy()
    for _ in range(3):
        o.exec([want], O.M_PLAIN, 1, O.D_DOUBLE, O.O_ADD, 0)
    u = _dev(u0)
    X.gs(u, X.gs_double, X.gs_add, 0, g)          # warm-up outside the capture (module load)
    u.copy_(torch.from_numpy(u0))
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_add, 0, g, s)
    u.copy_(torch.from_numpy(u0))                 # capture does not execute; replay does
    graph.replay()
    torch.cuda.synchronize()
    assert np.array_equal(u.cpu().numpy(), want)
    X.gs_free(g)
