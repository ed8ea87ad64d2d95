"""Pins the CPU oracle (oracle/gs_oracle.c) to what the reference itself
produces: the known answers recorded in SURVEY.md 3.4 / 8c (obtained by
running the reference's code), the assertions of the reference's own
tests/gs/gs_test.upc:118-240, and the golden fixtures under tests/golden/
generated from the compiled reference sources by oracle/make_golden.py."""
import glob
import itertools
import os

import numpy as np
import pytest

import cases
from exags_b200 import semmesh

NIL = 0xFFFFFFFF


def _ref_ok():
    import ref
    return ref.available()


def sent(*xs):
    return np.array([NIL if x < 0 else x for x in xs], dtype=np.uint32)


def test_known_answer_maps(O):
    o = O.Oracle([cases.known_answer_ids()])
    assert np.array_equal(o.map(0, 0), sent(2, 3, 4, -1, 5, 7, -1, 6, 9, -1, -1))
    assert np.array_equal(o.map(0, 1), sent(2, 3, 4, 1, -1, 5, 7, -1, 6, 9, 8, -1, 11, 12, 13, -1, 14, 15, -1, -1))
    assert np.array_equal(o.map(0, 2), sent(0, -1))


def test_known_answer_values(O):
    o = O.Oracle([cases.known_answer_ids()])
    v = np.arange(1, 17, dtype=np.float64)
    o.exec([v], O.M_PLAIN, 1, O.D_DOUBLE, O.O_ADD, 0)
    assert v.tolist() == [0, 12, 12, 12, 12, 14, 17, 14, 17, 17, 11, 12, 12, 12, 15, 15]
    v = np.arange(1, 17, dtype=np.float64)
    o.exec([v], O.M_PLAIN, 1, O.D_DOUBLE, O.O_ADD, 1)
    assert v.tolist() == [1, 2, 14, 14, 14, 14, 26, 14, 9, 26, 11, 39, 13, 14, 31, 16]
    v = np.arange(1, 17, dtype=np.float64)
    o.exec([v], O.M_PLAIN, 1, O.D_DOUBLE, O.O_MAX, 0)
    assert v.tolist() == [-np.finfo(np.float64).max, 5, 5, 5, 5, 8, 10, 8, 10, 10, 11, 12, 12, 12, 15, 15]


@pytest.mark.parametrize("np_", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("method", ["pairwise", "allreduce"])
def test_reference_gs_test_fixture(O, np_, method):
    """tests/gs/gs_test.upc:138-144: v[np+3]==3 (t=0) and ==np+3 (t=1)."""
    m = O.METHOD_PAIRWISE if method == "pairwise" else O.METHOD_ALLREDUCE
    ids = [cases.gs_test_fixture_ids(np_, r) for r in range(np_)]
    o = O.Oracle(ids, method=m)
    for t, want in ((0, 3.0), (1, float(np_ + 3))):
        vs = [np.ones(np_ + 4) for _ in range(np_)]
        o.exec(vs, O.M_PLAIN, 1, O.D_DOUBLE, O.O_ADD, t)
        for r in range(np_):
            assert vs[r][np_ + 3] == want


def test_small_sem_mesh(O):
    """2x2x2 hex, N=3 (BASELINE config 1): 512 DOF, 343 unique, sum of
    multiplicities 1000, max 8; gs(max,int) of iv[i]=i gives iv[3]=64 and
    gs(min,int) gives iv[3]=3."""
    ids = semmesh.box_ids(2, 2, 2, 3)
    assert ids.size == 512 and np.unique(ids).size == 343
    o = O.Oracle([ids])
    v = np.ones(512)
    o.exec([v], O.M_PLAIN, 1, O.D_DOUBLE, O.O_ADD, 0)
    assert v.sum() == 1000 and v.max() == 8
    iv = np.arange(512, dtype=np.int32)
    o.exec([iv], O.M_PLAIN, 1, O.D_INT, O.O_MAX, 0)
    assert iv[3] == 64
    iv = np.arange(512, dtype=np.int32)
    o.exec([iv], O.M_PLAIN, 1, O.D_INT, O.O_MIN, 0)
    assert iv[3] == 3


def test_partition_independence(O):
    """Concatenating all ranks' labels and running np=1 must give every copy
    the same result as the np-rank pairwise run (exact for ints, min/max)."""
    P = 3
    ids = [semmesh.box_ids(3, 2, 2, 3, ez0=2 * r, Ez_global=2 * P) for r in range(P)]
    vals = [cases.values(a.size, np.int64, "add", 10 + r) for r, a in enumerate(ids)]
    multi = O.Oracle(ids)
    vm = [v.copy() for v in vals]
    multi.exec(vm, O.M_PLAIN, 1, O.D_LONG, O.O_ADD, 0)
    one = O.Oracle([np.concatenate(ids)])
    v1 = np.concatenate(vals)
    one.exec([v1], O.M_PLAIN, 1, O.D_LONG, O.O_ADD, 0)
    assert np.array_equal(np.concatenate(vm), v1)


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_from_reference(O, path):
    """Fixtures written by oracle/make_golden.py from the reference's own
    compiled sources (oracle/_ref): inputs, maps and outputs."""
    z = np.load(path)
    np_ = int(z["np"])
    ids = [z[f"id{r}"] for r in range(np_)]
    method = int(z["method"])
    o = O.Oracle(ids, unique=int(z["unique"]),
                 method=O.METHOD_ALLREDUCE if method == 3 else O.METHOD_PAIRWISE)
    if "map0_0" in z:
        for r in range(np_):
            for w in (0, 1, 2):
                assert np.array_equal(o.map(r, w), z[f"map{w}_{r}"]), (path, r, w)
    k = 0
    while f"case{k}_meta" in z:
        mode, vn, dom, op, t = [int(x) for x in z[f"case{k}_meta"]]
        us = []
        for r in range(np_):
            a = z[f"case{k}_in{r}"].copy()
            us.append([a[i].copy() for i in range(vn)] if mode == O.M_MANY else a)
        o.exec(us, mode, vn, dom, op, t)
        for r in range(np_):
            got = np.stack(us[r]) if mode == O.M_MANY else us[r]
            want = z[f"case{k}_out{r}"]
            if want.dtype.kind == "f" and op in (O.O_ADD, O.O_MUL) and method == 3:
                # all-reduce association order differs from rank order: 4 ulp
                np.testing.assert_array_max_ulp(got, want, maxulp=4)
            else:
                assert np.array_equal(got, want), (path, k, r)
        k += 1


@pytest.mark.skipif(not _ref_ok(), reason="oracle/_ref (the reference's compiled sources) not built")
@pytest.mark.parametrize("name", list(__import__("local_cases").ENTRIES))
def test_local_entry_points_against_reference(O, name):
    """ogs_local (the oracle's restatement of the 14 public gs_local.h entry
    points, src/gs_local.h:23-41) against the reference's own compiled
    gs_local.c, every domain and op, bit for bit."""
    import copy
    import local_cases as LC
    import ref
    table = LC.maps(O)
    for (dname, dt), op in itertools.product(cases.DTYPES.items(), cases.OPS_BPR):
        if not LC.ENTRIES[name][4] and op != "add":
            continue
        out, inp, vn, mp = LC.make(name, table, dt, op)
        a_out = copy.deepcopy(out)
        a_in = a_out if inp is out else copy.deepcopy(inp)
        b_out = copy.deepcopy(out)
        b_in = b_out if inp is out else copy.deepcopy(inp)
        dom, opi = O.NP_DOM[np.dtype(dt)], cases.OPS_BPR.index(op)
        O.local(name, a_out, a_in, vn, mp, dom, opi)
        ref.local(name, b_out, b_in, vn, mp, dom, opi)
        same = all(np.array_equal(x, y) for x, y in zip(a_out, b_out)) if isinstance(out, list) else np.array_equal(a_out, b_out)
        assert same, f"{name} {dname} {op}"
