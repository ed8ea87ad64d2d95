"""Synthetic Nek5000-style spectral-element numbering and partition-independent
field values.  Harness code shared by tests/ and bench.py; it
feeds gs_setup, it is not part of the operator.

Box of Ex x Ey x Ez hexes of order N.  Element e = (ex,ey,ez), x fastest; node
(a,b,c) inside the element, a fastest; local index
    i = e*(N+1)^3 + a + (N+1)*(b + (N+1)*c)
global id (1-based, all positive => symmetric gs)
    id = 1 + gx + Gx*(gy + Gy*gz),  gx = ex*N + a, ...,  Gx = Ex*N + 1.
"""
from __future__ import annotations

import numpy as np

SEED = 20240607


def box_ids(Ex: int, Ey: int, Ez: int, N: int, ez0: int = 0, Ez_global: int | None = None,
            dtype=np.int64) -> np.ndarray:
    """Global ids of the elements with ez in [ez0, ez0+Ez) of an Ex x Ey x Ez_global box."""
    if Ez_global is None:
        Ez_global = Ez
    m = N + 1
    Gx, Gy = Ex * N + 1, Ey * N + 1
    gx = (np.arange(Ex, dtype=np.int64)[:, None] * N + np.arange(m, dtype=np.int64)[None, :])      # [Ex,a]
    gy = (np.arange(Ey, dtype=np.int64)[:, None] * N + np.arange(m, dtype=np.int64)[None, :])      # [Ey,b]
    gz = ((ez0 + np.arange(Ez, dtype=np.int64))[:, None] * N + np.arange(m, dtype=np.int64)[None, :])  # [Ez,c]
    # layout [ez, ey, ex, c, b, a]
    ids = (1 + gx[None, None, :, None, None, :]
           + Gx * (gy[None, :, None, None, :, None]
                   + Gy * gz[:, None, None, :, None, None]))
    return np.ascontiguousarray(ids.reshape(-1)).astype(dtype, copy=False)


def box_counts(Ex: int, Ey: int, Ez: int, N: int) -> dict:
    """Analytic n, unique nodes, groups G and member indices S of a whole box."""
    m = N + 1
    n = Ex * Ey * Ez * m ** 3
    G = [E * N + 1 for E in (Ex, Ey, Ez)]
    U = [g - (E - 1) for g, E in zip(G, (Ex, Ey, Ez))]
    return {"n": n, "unique": G[0] * G[1] * G[2],
            "groups": G[0] * G[1] * G[2] - U[0] * U[1] * U[2],
            "members": n - U[0] * U[1] * U[2]}


def block_partition_ids(Ex: int, Ey: int, Ez: int, N: int, bx: int, by: int, bz: int,
                        nranks: int, rank: int, seed: int = 12345) -> tuple[np.ndarray, np.ndarray]:
    """Config-5 style irregular partition: elements grouped in bx x by x bz
    blocks, block b -> rank perm[b] % nranks with perm a fixed LCG shuffle.
    Returns (ids of this rank's elements, their global element numbers)."""
    nbx, nby, nbz = Ex // bx, Ey // by, Ez // bz
    nblk = nbx * nby * nbz
    perm = np.arange(nblk, dtype=np.int64)
    state = seed
    for k in range(nblk - 1, 0, -1):           # Fisher-Yates driven by an LCG
        state = (state * 1103515245 + 12345) & 0x7FFFFFFF
        j = state % (k + 1)
        perm[k], perm[j] = perm[j], perm[k]
    ex, ey, ez = np.meshgrid(np.arange(Ex), np.arange(Ey), np.arange(Ez), indexing="ij")
    blk = (ex // bx) + nbx * ((ey // by) + nby * (ez // bz))
    owner = perm[blk] % nranks
    eglob = (ex + Ex * (ey + Ey * ez))
    mine = np.sort(eglob[owner == rank].ravel())
    m = N + 1
    Gx, Gy = Ex * N + 1, Ey * N + 1
    e_x, e_y, e_z = mine % Ex, (mine // Ex) % Ey, mine // (Ex * Ey)
    a = np.arange(m, dtype=np.int64)
    gx = e_x[:, None, None, None] * N + a[None, None, None, :]
    gy = e_y[:, None, None, None] * N + a[None, None, :, None]
    gz = e_z[:, None, None, None] * N + a[None, :, None, None]
    ids = 1 + gx + Gx * (gy + Gy * gz)
    return np.ascontiguousarray(ids.reshape(-1)), mine


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def field_values(dof_global: np.ndarray, dtype, op: str = "add", k: int = 0,
                 seed: int = SEED) -> np.ndarray:
    """Value of each DOF from a hash of its global DOF number (element*(N+1)^3
    + node), so the field does not depend on the partition.
    FP: uniform [0.5,1.5); int add/min/max: [-1000,1000]; int mul: {1,2}."""
    with np.errstate(over="ignore"):
        h = _splitmix64((dof_global.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))
                        ^ np.uint64(seed) ^ (np.uint64(k) << np.uint64(56)))
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        return (0.5 + (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))).astype(dtype)
    if op == "mul":
        return (1 + (h & np.uint64(1))).astype(dtype)
    return ((h % np.uint64(2001)).astype(np.int64) - 1000).astype(dtype)


def slab_dof_numbers(Ex: int, Ey: int, Ez: int, N: int, ez0: int = 0) -> np.ndarray:
    """Global DOF numbers (element_global*(N+1)^3 + node) of a z-slab."""
    m3 = (N + 1) ** 3
    first = ez0 * Ex * Ey * m3
    return np.arange(first, first + Ex * Ey * Ez * m3, dtype=np.uint64)
