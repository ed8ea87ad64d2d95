"""Index logic of the device gs_setup passes (csrc/setup_pipeline.h), checked
without a GPU: tests/c/setup_emul.cpp instantiates the same pass source with a
sequential host backend and compares every array (topology rows, local map,
flagged-primary lists, sector count, sliced-ELL layout) with the host builders
of csrc/topo.c.  The CUDA backend itself is covered by the -m gpu tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cases
from exags_b200 import semmesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "exaflow-exags-upc_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "c", "build")


@pytest.fixture(scope="module")
def emul():
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for c in ("topo.c", "sort.c", "base.c"):
        o = os.path.join(OUT, c[:-2] + ".o")
        subprocess.run(["gcc", "-O2", "-fPIC", "-fopenmp", "-std=gnu11", "-I", os.path.join(ROOT, "include"),
                        "-c", os.path.join(CSRC, c), "-o", o], check=True)
        objs.append(o)
    so = os.path.join(OUT, "libsetup_emul.so")
    subprocess.run(["g++", "-O2", "-fPIC", "-std=c++17", "-shared", "-fopenmp", "-I", CSRC,
                    os.path.join(ROOT, "tests", "c", "setup_emul.cpp")] + objs + ["-o", so], check=True)
    L = ctypes.CDLL(so)
    L.setup_emul_compare.restype = ctypes.c_int
    L.setup_emul_compare.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_int]

    L.setup_emul_split.restype = ctypes.c_int
    L.setup_emul_split.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p,
                                   ctypes.c_uint, ctypes.c_int]
    L.setup_emul_ngroups.restype = ctypes.c_size_t
    L.setup_emul_ngroups.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]

    def split(ids, frac, seed, reverse=0):
        ids = np.ascontiguousarray(ids)
        ng = L.setup_emul_ngroups(ids.ctypes.data, ids.dtype.itemsize, ids.size)
        rng = np.random.default_rng(seed)
        bgrp = np.flatnonzero(rng.random(ng) < frac).astype(np.uint32)
        return L.setup_emul_split(ids.ctypes.data, ids.dtype.itemsize, ids.size, bgrp.ctypes.data, bgrp.size, reverse)

    L.setup_emul_candidates.restype = ctypes.c_int
    L.setup_emul_candidates.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.c_int, ctypes.c_int]
    L.setup_emul_sell_roundtrip.restype = ctypes.c_int
    L.setup_emul_sell_roundtrip.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
    L.setup_emul_line_visits.restype = ctypes.c_ulonglong
    L.setup_emul_line_visits.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]

    def roundtrip(ids):
        ids = np.ascontiguousarray(ids)
        return L.setup_emul_sell_roundtrip(ids.ctypes.data, ids.dtype.itemsize, ids.size)

    def visits(ids):
        ids = np.ascontiguousarray(ids)
        return L.setup_emul_line_visits(ids.ctypes.data, ids.dtype.itemsize, ids.size)

    def candidates(a, b, reverse=0):
        a, b = np.ascontiguousarray(a, dtype=np.int64), np.ascontiguousarray(b, dtype=np.int64)
        return L.setup_emul_candidates(a.ctypes.data, a.size, b.ctypes.data, b.size, 8, reverse)

    def run(ids, reverse=0):
        ids = np.ascontiguousarray(ids)
        assert ids.dtype in (np.int64, np.int32)
        return L.setup_emul_compare(ids.ctypes.data, ids.dtype.itemsize, ids.size, reverse)
    run.split = split
    run.roundtrip = roundtrip
    run.visits = visits
    run.candidates = candidates
    return run


@pytest.mark.parametrize("seed", range(8))
def test_random_labels(emul, seed):
    ids = cases.random_ids(3000 + 211 * seed, 700, seed, long_group=40 if seed % 2 else 0)
    assert emul(ids) == 0
    assert emul(ids, reverse=1) == 0


@pytest.mark.parametrize("seed", range(3))
def test_random_labels_unflagged(emul, seed):
    ids = cases.random_ids(5000, 900, 100 + seed, p_flag=0.0, long_group=17 if seed else 0)
    assert emul(ids) == 0
    assert emul(ids.astype(np.int32), reverse=1) == 0


@pytest.mark.parametrize("ids", [np.zeros(0, np.int64), np.zeros(5, np.int64),
                                 np.array([7], np.int64), np.array([-7], np.int64),
                                 np.arange(1, 50, dtype=np.int64),
                                 np.array([3, 3], np.int64), np.array([-3, 3, 0, -3], np.int64),
                                 np.full(9, 5, np.int64), np.full(8, -5, np.int64)],
                         ids=["empty", "allzero", "single", "singleflag", "nogroups", "pair", "flagpair",
                              "nine", "eightflagged"])
def test_degenerate(emul, ids):
    assert emul(ids) == 0


@pytest.mark.parametrize("dims", [(2, 2, 2, 3), (5, 4, 3, 4), (16, 16, 6, 7), (7, 5, 3, 9), (40, 3, 2, 1)])
def test_sem_boxes(emul, dims):
    """Several windows of the sliced layout (16x16x6 N=7 has ~380 of them)."""
    ids = semmesh.box_ids(*dims)
    assert emul(ids) == 0
    assert emul(ids, reverse=1) == 0


def test_window_boundaries(emul):
    """Group sizes chosen so that windows end with unused slots (a 4- or 8-wide
    group that does not fit the remaining 2 or 6 slots) and lanes are skipped."""
    rng = np.random.default_rng(5)
    sizes = rng.choice([2, 3, 5, 8, 2, 2, 7], size=6000)
    labels = np.repeat(np.arange(1, sizes.size + 1), sizes)
    ids = labels[rng.permutation(labels.size)].astype(np.int64)
    assert emul(ids) == 0
    # primaries in label order: whole windows of one width, then a mix
    sizes = np.concatenate([np.full(1500, 2), np.full(700, 8), np.full(900, 4), rng.choice([2, 4, 8], 3000)])
    ids = np.repeat(np.arange(1, sizes.size + 1), sizes).astype(np.int64)
    assert emul(ids) == 0
    assert emul(ids, reverse=1) == 0


def test_large_labels(emul):
    ids = np.array([2**61, 5, -(2**61), 2**40, 5, 2**40, 0, 1], np.int64)
    assert emul(ids) == 0


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("frac", [0.0, 0.2, 1.0])
def test_interior_boundary_split(emul, seed, frac):
    """A random subset of the groups plays the boundary an exchange plan names."""
    ids = cases.random_ids(2500 + 97 * seed, 600, 40 + seed, p_flag=0.0 if seed % 2 else 0.25,
                           long_group=30 if seed % 3 == 0 else 0)
    assert emul.split(ids, frac, seed) == 0
    assert emul.split(ids, frac, seed, reverse=1) == 0


def test_interior_boundary_split_mesh(emul):
    ids = semmesh.box_ids(6, 5, 4, 5)
    assert emul.split(ids, 0.1, 1) == 0
    assert emul.split(np.zeros(0, np.int64), 0.5, 2) == 0


@pytest.mark.parametrize("pairs", ["2", "1", "0"])
def test_sliced_layout_decodes_to_the_map(emul, pairs, monkeypatch):
    """The layout (with and without the pair ordering) holds every group once,
    members in order, and phase bits only on pairs."""
    monkeypatch.setenv("EXAGS_SELL_PAIRS", pairs)
    for seed in range(4):
        ids = cases.random_ids(4000 + 301 * seed, 900, 60 + seed, long_group=25 if seed % 2 else 0)
        assert emul.roundtrip(ids) == 0
    for dims in [(2, 2, 2, 3), (16, 16, 6, 7), (7, 5, 3, 9), (40, 1, 1, 7)]:
        assert emul.roundtrip(semmesh.box_ids(*dims)) == 0


def test_pair_ordering_reduces_line_visits(emul, monkeypatch):
    """x-face pairs of a row of elements: about half the 128-byte line visits
    per load instruction (64 -> 34 per interface; 32 is the floor)."""
    ids = semmesh.box_ids(64, 1, 1, 7)
    monkeypatch.setenv("EXAGS_SELL_PAIRS", "0")
    plain = emul.visits(ids)
    monkeypatch.setenv("EXAGS_SELL_PAIRS", "1")
    chained = emul.visits(ids)
    assert chained < 0.6 * plain, (plain, chained)
    ids = semmesh.box_ids(12, 12, 4, 7)
    monkeypatch.setenv("EXAGS_SELL_PAIRS", "0")
    plain = emul.visits(ids)
    monkeypatch.setenv("EXAGS_SELL_PAIRS", "1")
    pairs_only = emul.visits(ids)
    monkeypatch.setenv("EXAGS_SELL_PAIRS", "2")
    assert emul.visits(ids) < pairs_only < 0.92 * plain


@pytest.mark.parametrize("reverse", [0, 1])
def test_candidate_filters_never_drop_a_shared_label(emul, reverse):
    """The device passes that pick, with np > 1, the labels another rank may hold
    (setup_pipeline.h; the filters in front of shared_ids, src/gs.upc:190-244):
    two label sets stand in for two ranks.  A label both hold must be a candidate,
    the candidate rows must be the groups' (|label|, primary) in group order, and the
    group lookup of the exchange plan's primaries must find every primary again."""
    rng = np.random.default_rng(5)
    slabs = [semmesh.box_ids(4, 3, 2, 3, ez0=2 * r, Ez_global=4) for r in range(2)]
    assert emul.candidates(slabs[0], slabs[1], reverse) == 0
    assert emul.candidates(slabs[1], slabs[0], reverse) == 0
    for seed in range(6):
        a = cases.random_ids(900 + 50 * seed, 300, 40 + seed)
        b = cases.random_ids(700, 300, 80 + seed)
        assert emul.candidates(a, b, reverse) == 0, seed
    big = (rng.choice(2**20, size=500, replace=False).astype(np.int64) << 38) + 7     # widely spaced labels
    assert emul.candidates(big[:300], big[200:], reverse) == 0
    assert emul.candidates(np.arange(1, 100), np.arange(1000, 1100), reverse) == 0    # nothing shared
    assert emul.candidates(np.zeros(0, np.int64), np.arange(1, 10), reverse) == 0
