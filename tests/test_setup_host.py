"""Host-side gs_setup logic against the oracle, without a GPU: local maps at
np=1 and, through 2..4 real processes (gloo for the test's bookkeeping, the
library's own shared-memory rendezvous for its id exchange), the shared-label
discovery, pairwise maps and gs_unique."""
import numpy as np
import pytest

import cases
from exags_b200 import semmesh
from mp_launch import run_ranks


@pytest.mark.parametrize("seed", range(6))
def test_local_maps_random(X, O, seed):
    ids = cases.random_ids(2000 + 137 * seed, 500, seed, long_group=40 if seed % 2 else 0)
    g = X.gs_setup(ids)
    o = O.Oracle([ids])
    for w in (0, 1, 2):
        assert np.array_equal(g.dump_map(w), o.map(0, w)), w
    X.gs_free(g)


@pytest.mark.parametrize("ids", [np.zeros(0, np.int64), np.zeros(5, np.int64),
                                 np.array([7], np.int64), np.array([-7], np.int64),
                                 np.arange(1, 50, dtype=np.int64)],
                         ids=["empty", "allzero", "single", "singleflag", "nogroups"])
def test_local_maps_degenerate(X, O, ids):
    g = X.gs_setup(ids)
    o = O.Oracle([ids])
    for w in (0, 1, 2):
        assert np.array_equal(g.dump_map(w), o.map(0, w)), w
    X.gs_free(g)


def test_int32_flavour_matches_int64(X):
    ids = cases.random_ids(900, 200, 3)
    a, b = X.gs_setup(ids), X.gs_setup(ids.astype(np.int32))
    for w in (0, 1, 2):
        assert np.array_equal(a.dump_map(w), b.dump_map(w))
    X.gs_free(a); X.gs_free(b)


def test_sem_mesh_counts(X):
    """Group/member counts of a box match the closed forms of SURVEY.md 8."""
    dims = (5, 4, 3, 4)
    g = X.gs_setup(semmesh.box_ids(*dims))
    c = semmesh.box_counts(*dims)
    info = g.info()
    assert (info["n"], info["groups"], info["members"]) == (c["n"], c["groups"], c["members"])
    X.gs_free(g)


def test_unique_np1(X, O):
    ids = cases.random_ids(1500, 300, 11, p_flag=0.0)
    mine, want = ids.copy(), [ids.copy()]
    X.gs_unique(mine)
    O.unique(want)
    assert np.array_equal(mine, want[0])
    g, o = X.gs_setup(ids, unique=1), O.Oracle([ids], unique=1)
    for w in (0, 1, 2):
        assert np.array_equal(g.dump_map(w), o.map(0, w)), w
    X.gs_free(g)


@pytest.mark.parametrize("np_,case", [(2, "fixture"), (2, "slab"), (3, "random"), (4, "blocks"),
                                      (2, "empty"), (4, "fixture"), (8, "blocks"), (8, "fixture"),
                                      (5, "random"), (4, "bigslab"), (3, "bigblocks"),
                                      (2, "sparse"), (3, "sparse"), (4, "clustered"), (2, "clustered")])
def test_multirank_setup(built, np_, case):
    codes, outs = run_ranks(np_, [case, "setup"])
    assert all(c == 0 for c in codes), "\n".join(outs)


@pytest.mark.parametrize("np_,case", [(2, "slab"), (3, "blocks")])
def test_multirank_setup_unique(built, np_, case):
    codes, outs = run_ranks(np_, [case, "setup", 1])
    assert all(c == 0 for c in codes), "\n".join(outs)


@pytest.mark.parametrize("np_,case", [(2, "slab"), (3, "random"), (4, "blocks")])
def test_multirank_gs_unique(built, np_, case):
    codes, outs = run_ranks(np_, [case, "unique"])
    assert all(c == 0 for c in codes), "\n".join(outs)


@pytest.mark.parametrize("np_", [2, 3, 5])
def test_multirank_comm_collectives(built, np_):
    """comm_allreduce (every dom x op; vectors small enough for the control
    segment and large enough for published blocks), comm_scan, comm_dot,
    comm_bcast and comm_reduce on host vectors across ranks
    (src/comm.upc:427-459,706-984): folded in rank order on every rank."""
    codes, outs = run_ranks(np_, ["fixture", "comm"])
    assert all(c == 0 for c in codes), "\n".join(outs)


@pytest.mark.parametrize("np_,case", [(2, "slab"), (4, "blocks")])
def test_multirank_setup_crystal_router_request(built, np_, case):
    """gs_crystal_router requests (src/gs.h:127) run the direct pairwise
    exchange: the handle must carry the pairwise maps."""
    codes, outs = run_ranks(np_, [case, "setup", 0, 2])
    assert all(c == 0 for c in codes), "\n".join(outs)


@pytest.mark.parametrize("np_,seed", [(2, 1), (3, 2), (4, 3), (7, 6)])
def test_multirank_setup_fuzz(built, np_, seed):
    """Randomised label sets (sizes, density, zeros, flags, a long group drawn
    from the seed): all five maps of every rank against the oracle, with and
    without unique."""
    for unique in (0, 1):
        codes, outs = run_ranks(np_, [f"fuzz{seed}", "setup", unique])
        assert all(c == 0 for c in codes), "\n".join(outs)


def test_dead_rank_ends_the_wait(built):
    """A rank that dies while the others sit in a host barrier: they fail through
    the reference's fail() contract (src/fail.upc:19-63) within seconds, not hang."""
    import time
    t0 = time.time()
    codes, outs = run_ranks(3, ["fixture", "die", 0, 1], timeout=120)
    assert time.time() - t0 < 60
    assert codes[2] == 7
    assert codes[0] == 1 and codes[1] == 1, (codes, outs)
    assert "is gone while rank" in outs[0] and "ERROR (proc" in outs[0]


def test_stale_rendezvous_segment_is_replaced(built):
    """A rendezvous segment left in /dev/shm by a killed run with the same job key
    (only atexit unlinks it) carries a dead owner: late ranks must not attach to
    it, rank 0 replaces it, and the job runs (comm.c rendezvous)."""
    import ctypes
    import mp_launch
    import uuid
    key = "stale" + uuid.uuid4().hex[:10]
    path = f"/dev/shm/exags-{key}-ctl"
    # what a dead run leaves behind: right size, magic set, every rank "attached", owner pid gone
    XW_MAXNP, XW_SMALL = 64, 4096
    size = 4 * 5 + 4 * XW_MAXNP + 8 + 8 * XW_MAXNP + XW_MAXNP * XW_SMALL + 64
    blob = bytearray(size)
    blob[0:4] = (0x45584147).to_bytes(4, "little")      # magic
    blob[4:8] = (2).to_bytes(4, "little")               # np
    blob[8:12] = (2).to_bytes(4, "little")              # attached
    blob[20:24] = (2 ** 22 - 3).to_bytes(4, "little")   # pid[0]: no such process
    blob[280:288] = (12345).to_bytes(8, "little")       # owner_start (after pid[64], 8-byte aligned)
    with open(path, "wb") as f:
        f.write(blob)
    try:
        port = 20000 + (uuid.uuid4().int % 30000)
        procs = []
        import os, subprocess, sys
        for r in range(2):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                       MASTER_PORT=str(port), EXAGS_JOB_KEY=key, OMP_NUM_THREADS="2",
                       TEST_RANK0_DELAY_S="2")      # rank 1 meets the stale segment before rank 0 replaces it
            procs.append(subprocess.Popen([sys.executable, os.path.join(mp_launch.HERE, "mp_worker.py"),
                                           "fixture", "setup", "0", "1"], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=180)[0] for p in procs]
        assert [p.returncode for p in procs] == [0, 0], outs
    finally:
        if os.path.exists(path):
            os.unlink(path)
