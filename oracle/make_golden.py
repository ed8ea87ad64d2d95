#!/usr/bin/env python3
"""make_golden.py -- write tests/golden/*.npz from the REFERENCE's own code.

Runs oracle/_ref/libexags_ref.so (built by oracle/build_ref.sh from
/root/reference/src) on small seeded jobs and stores inputs, the reference's
three local maps and its outputs.  The fixtures are committed; the reference
tree and the .so are not needed to run the tests that read them
(tests/test_oracle_pin.py::test_golden_from_reference and the GPU parity
tests).  Re-run after `make -C oracle ref`:   python oracle/make_golden.py
"""
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(ROOT, "tests"), os.path.join(ROOT, "exaflow-exags-upc_b200")]
import cases  # noqa: E402
import ref  # noqa: E402
from exags_b200 import semmesh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DT = [np.float64, np.float32, np.int32, np.int64]
OPS = ["add", "mul", "min", "max"]


def job(name, ids, combos, unique=0, method=1, seed=0):
    """combos: list of (mode, vn, dom, op, t)."""
    np_ = len(ids)
    if np_ > 1:
        # The reference grows its pairwise mailboxes inside pw_exec with a
        # collective (barriers in comm_alloc_thrd_buf, src/comm.upc:139-184) guarded
        # by a rank-local size test (src/gs.upc:487): ranks whose buffers are
        # already big enough skip the barriers and the job deadlocks.  A first,
        # widest case sizes every rank's mailboxes once.
        combos = [(1, 8, 0, 0, 0)] + list(combos)
    data = []
    for k, (mode, vn, dom, op, t) in enumerate(combos):
        row = []
        for r in range(np_):
            nv = 1 if mode == 0 else vn
            row.append(cases.values(ids[r].size * nv, DT[dom], OPS[op], seed + 97 * k + r))
        data.append(row)
    inputs = [[a.copy() for a in row] for row in data]
    maps, _ = ref.run(ids, combos, data, unique=unique, method=method, want_maps=True)
    z = {"np": np.int64(np_), "unique": np.int64(unique), "method": np.int64(method)}
    for r in range(np_):
        z[f"id{r}"] = np.asarray(ids[r], dtype=np.int64)
        for w in range(3):
            z[f"map{w}_{r}"] = maps[r][w]
    for k, (mode, vn, dom, op, t) in enumerate(combos):
        z[f"case{k}_meta"] = np.array([mode, vn, dom, op, t], dtype=np.int64)
        for r in range(np_):
            n = ids[r].size
            shape = (vn, n) if mode == 2 else (-1,)
            z[f"case{k}_in{r}"] = inputs[k][r].reshape(shape)
            z[f"case{k}_out{r}"] = data[k][r].reshape(shape)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **z)
    print("wrote", name, "ranks", np_, "cases", len(combos))


def main():
    if not ref.available():
        raise SystemExit("oracle/_ref/libexags_ref.so missing: run `make -C oracle ref` first")
    plain_all = [(0, 1, d, o, t) for d, o, t in itertools.product(range(4), range(4), (0, 1))]
    modes = [(0, 1), (1, 3), (2, 3)]
    mixed = [(m, vn, d, o, t) for (m, vn), (d, o), t in
             itertools.product(modes, ((0, 0), (1, 1), (2, 2), (3, 3), (0, 3), (3, 0)), (0, 1))]
    job("g01_known_np1", [cases.known_answer_ids()], plain_all)
    for p in (1, 2, 4):
        job(f"g02_fixture_np{p}", [cases.gs_test_fixture_ids(p, r) for r in range(p)],
            [(0, 1, 0, 0, 0), (0, 1, 0, 0, 1), (0, 1, 2, 3, 0), (1, 2, 3, 0, 1)])
    job("g03_random_np1", [cases.random_ids(400, 90, 3, long_group=20)], mixed, seed=5)
    # powers of two only: the reference tests gs on 1/2/4/8/16 threads
    # (tests/gs/Makefile.am:1-11) and its hand-rolled collectives do not
    # terminate on 3 threads under the shim
    job("g04_random_np4", [cases.random_ids(300 + 17 * r, 110, 40 + r) for r in range(4)], mixed, seed=9)
    job("g05_slab_np2", [semmesh.box_ids(3, 3, 2, 3, ez0=2 * r, Ez_global=4) for r in range(2)],
        [(0, 1, 0, 0, 0), (1, 3, 0, 0, 0), (2, 3, 1, 0, 0), (0, 1, 2, 2, 0), (0, 1, 3, 3, 1)], seed=11)
    job("g06_blocks_np4_pairwise", [semmesh.block_partition_ids(4, 4, 4, 2, 2, 2, 2, 4, r)[0] for r in range(4)],
        [(0, 1, 0, 0, 0), (0, 1, 3, 0, 0), (1, 3, 0, 1, 0), (2, 3, 2, 3, 1)], seed=13)
    job("g07_blocks_np4_allreduce", [semmesh.block_partition_ids(4, 4, 4, 2, 2, 2, 2, 4, r)[0] for r in range(4)],
        [(0, 1, 0, 0, 0), (0, 1, 3, 0, 0), (0, 1, 2, 2, 1), (1, 3, 0, 3, 0)], method=3, seed=13)
    job("g08_unique_np2", [semmesh.box_ids(3, 3, 2, 3, ez0=2 * r, Ez_global=4) for r in range(2)],
        [(0, 1, 0, 0, 0), (0, 1, 0, 0, 1), (0, 1, 3, 3, 0)], unique=1, seed=17)
    job("g09_config1_2x2x2_N3", [semmesh.box_ids(2, 2, 2, 3)],
        [(0, 1, d, o, 0) for d, o in itertools.product((0, 2), (0, 2, 3))], seed=19)
    job("g10_flagged_np2_allreduce", [cases.random_ids(250, 60, 70 + r) for r in range(2)],
        [(0, 1, 0, 0, 0), (0, 1, 0, 0, 1), (0, 1, 3, 3, 0), (0, 1, 2, 2, 1)], method=3, seed=23)


if __name__ == "__main__":
    main()
