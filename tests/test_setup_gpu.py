"""gs_setup on the device (csrc/setup_cuda.cu) against the host builders
(csrc/topo.c, EXAGS_SETUP=host): the same handle, array for array -- group
map, flagged primaries, sector count, sliced-layout size -- and the same gs
results, bit for bit."""
import os

import numpy as np
import pytest
import torch

import cases
from exags_b200 import semmesh

pytestmark = pytest.mark.gpu


def both(X, ids, **kw):
    old = os.environ.get("EXAGS_SETUP")
    try:
        os.environ["EXAGS_SETUP"] = "host"
        h = X.gs_setup(ids, **kw)
        os.environ["EXAGS_SETUP"] = "device"
        d = X.gs_setup(ids, **kw)
    finally:
        if old is None:
            os.environ.pop("EXAGS_SETUP", None)
        else:
            os.environ["EXAGS_SETUP"] = old
    return h, d


def same_handle(X, h, d):
    ih, idv = h.info(), d.info()
    assert ih == idv, (ih, idv)
    for w in (0, 1, 2):
        assert np.array_equal(h.dump_map(w), d.dump_map(w)), w


def same_results(X, h, d, n, seed=0):
    for dt, dom in ((np.float64, X.gs_double), (np.int32, X.gs_int)):
        for opname, op in (("add", X.gs_add), ("min", X.gs_min)):
            for tr in (0, 1):
                v = cases.values(n, dt, opname, seed)
                a = torch.from_numpy(v.copy()).cuda(); b = torch.from_numpy(v.copy()).cuda()
                X.gs(a, dom, op, tr, h); X.gs(b, dom, op, tr, d)
                assert torch.equal(a, b), (dt, opname, tr)
    v3 = cases.values(3 * n, np.float64, "add", seed + 1)
    a = torch.from_numpy(v3.copy()).cuda(); b = torch.from_numpy(v3.copy()).cuda()
    X.gs_vec(a, 3, X.gs_double, X.gs_add, 0, h); X.gs_vec(b, 3, X.gs_double, X.gs_add, 0, d)
    assert torch.equal(a, b)
    host = cases.values(n, np.float64, "add", seed + 2)
    ha, hb = host.copy(), host.copy()
    X.gs(ha, X.gs_double, X.gs_add, 0, h); X.gs(hb, X.gs_double, X.gs_add, 0, d)   # host pointers
    assert np.array_equal(ha, hb)


@pytest.mark.parametrize("seed", range(6))
def test_random_labels(X, seed):
    ids = cases.random_ids(3000 + 211 * seed, 700, seed, long_group=40 if seed % 2 else 0)
    h, d = both(X, ids)
    same_handle(X, h, d)
    same_results(X, h, d, ids.size, seed)
    X.gs_free(h); X.gs_free(d)


@pytest.mark.parametrize("ids", [np.zeros(0, np.int64), np.zeros(5, np.int64),
                                 np.array([7], np.int64), np.array([-7], np.int64),
                                 np.arange(1, 50, dtype=np.int64), np.array([-3, 3, 0, -3], np.int64),
                                 np.full(9, 5, np.int64), np.full(8, -5, np.int64)],
                         ids=["empty", "allzero", "single", "singleflag", "nogroups", "flagpair",
                              "nine", "eightflagged"])
def test_degenerate(X, ids):
    h, d = both(X, ids)
    same_handle(X, h, d)
    if ids.size:
        same_results(X, h, d, ids.size)
    X.gs_free(h); X.gs_free(d)


@pytest.mark.parametrize("dims", [(2, 2, 2, 3), (5, 4, 3, 4), (16, 16, 6, 7), (7, 5, 3, 9)])
def test_sem_boxes(X, dims):
    ids = semmesh.box_ids(*dims)
    h, d = both(X, ids)
    same_handle(X, h, d)
    same_results(X, h, d, ids.size)
    X.gs_free(h); X.gs_free(d)


def test_int32_labels_unique_and_device_resident_labels(X, O):
    ids = cases.random_ids(4000, 900, 21, p_flag=0.0)
    h, d = both(X, ids.astype(np.int32))
    same_handle(X, h, d)
    X.gs_free(h); X.gs_free(d)
    h, d = both(X, ids, unique=1)
    same_handle(X, h, d)
    o = O.Oracle([ids], unique=1)
    for w in (0, 1, 2):
        assert np.array_equal(d.dump_map(w), o.map(0, w)), w
    X.gs_free(h); X.gs_free(d)
    # labels already in HBM (extension): same handle as from host labels
    h = X.gs_setup(ids)
    for t in (torch.from_numpy(ids).cuda(), torch.from_numpy(ids.astype(np.int32)).cuda()):
        d = X.gs_setup(t)
        same_handle(X, h, d)
        X.gs_free(d)
    d = X.gs_setup(torch.from_numpy(ids).cuda(), unique=1)
    X.gs_free(d)
    X.gs_free(h)


def test_large_mesh_matches_host_and_is_faster(X):
    """12.6 M DOF box: identical handle, identical gs result, and the setup
    times of both builders printed (device setup is what gs_setup uses)."""
    import time
    ids = semmesh.box_ids(32, 32, 24, 7)
    t0 = time.perf_counter(); h, d = None, None
    os.environ["EXAGS_SETUP"] = "host"
    try:
        h = X.gs_setup(ids); t1 = time.perf_counter()
        os.environ["EXAGS_SETUP"] = "device"
        d = X.gs_setup(ids); t2 = time.perf_counter()
        d2 = X.gs_setup(ids); t3 = time.perf_counter()
    finally:
        os.environ.pop("EXAGS_SETUP", None)
    print(f"\nsetup n={ids.size}: host {t1 - t0:.3f} s, device {t2 - t1:.3f} s, device again {t3 - t2:.3f} s")
    assert h.info() == d.info()
    assert np.array_equal(h.dump_map(1), d.dump_map(1))
    v = cases.values(ids.size, np.float64, "add", 3)
    a = torch.from_numpy(v.copy()).cuda(); b = torch.from_numpy(v.copy()).cuda()
    X.gs(a, X.gs_double, X.gs_add, 0, h); X.gs(b, X.gs_double, X.gs_add, 0, d)
    assert torch.equal(a, b)
    X.gs_free(h); X.gs_free(d); X.gs_free(d2)
