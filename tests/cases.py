"""Shared test inputs: label sets with the edge cases the domain has
(zeros, flagged labels, singletons, long groups, empty input) and meshes."""
import numpy as np


def known_answer_ids():
    # SURVEY.md 3.4 -- outputs below were produced by running the reference
    return np.array([-11, -1, 1, 1, 1, 5, 7, 5, -7, 7, 0, 9, -9, -9, 3, -3], dtype=np.int64)


def random_ids(n, nlabels, seed, p_zero=0.1, p_flag=0.25, long_group=0):
    rng = np.random.default_rng(seed)
    ids = rng.integers(1, nlabels + 1, size=n).astype(np.int64)
    if long_group:
        ids[rng.choice(n, size=min(n, long_group), replace=False)] = nlabels + 7
    ids[rng.random(n) < p_zero] = 0
    flag = rng.random(n) < p_flag
    ids[flag] = -ids[flag]
    return ids


def gs_test_fixture_ids(np_, rank):
    """tests/gs/gs_test.upc:125-133"""
    ids = np.empty(np_ + 4, dtype=np.int64)
    ids[0] = -(np_ + 10 + 3 * rank)
    ids[1:np_ + 1] = -np.arange(1, np_ + 1)
    ids[np_ + 1] = rank + 1
    ids[np_ + 2] = rank + 1
    ids[np_ + 3] = np_ - rank
    return ids


def values(n, dtype, op, seed):
    rng = np.random.default_rng(seed)
    dtype = np.dtype(dtype)
    if op == "bpr":           # common binary prefix of unsigned values (src/gs_defs.h:37-42)
        return rng.integers(0, 1 << 20, size=n).astype(dtype)
    if dtype.kind == "f":
        return rng.uniform(0.5, 1.5, size=n).astype(dtype)
    if op == "mul":
        return rng.integers(1, 3, size=n).astype(dtype)
    return rng.integers(-1000, 1001, size=n).astype(dtype)


DTYPES = {"double": np.float64, "float": np.float32, "int": np.int32, "long": np.int64}
OPS = ["add", "mul", "min", "max"]
OPS_BPR = OPS + ["bpr"]     # gs_op order (src/gs_defs.h:73-79); bpr is what xxt.upc:623 issues
