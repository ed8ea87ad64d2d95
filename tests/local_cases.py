"""Inputs for the fourteen public gs_local.h entry points (src/gs_local.h:23-41)
on sentinel maps taken from the oracle's own setup: the local map of a random
label set (groups of many sizes, flagged and unlabelled entries) and the
pairwise receive / send maps of a 2-rank job ([index, slot, ..., -1]).  Shared
by the CPU pin test (oracle restatement against the reference's compiled
gs_local.c) and the GPU parity test (CUDA entry points against the oracle)."""
import numpy as np

import cases

NIL = 0xFFFFFFFF
VN = 3
NARR = 1000


def extents(mp):
    """(largest first entry + 1, largest following entry + 1) of a sentinel map"""
    first, rest, at_head = 0, 0, True
    for v in mp[:-1]:
        if v == NIL:
            at_head = True
        elif at_head:
            first, at_head = max(first, int(v) + 1), False
        else:
            rest = max(rest, int(v) + 1)
    return first, rest


def maps(O):
    ids = cases.random_ids(600, 140, 31, long_group=12)
    loc = O.Oracle([ids]).map(0, 1)
    two = [cases.random_ids(500 + 40 * r, 160, 77 + r, p_flag=0.1) for r in range(2)]
    o2 = O.Oracle(two)
    return {"local": (loc, ids.size), "recv": (o2.map(0, 3), two[0].size), "send": (o2.map(0, 4), two[0].size)}


# entry -> (map name, layout of out, layout of in, aliased, takes op)
#   layouts: "plain", "vec" ([i][VN]), "many" (VN arrays), None (no input)
ENTRIES = {
    "gs_gather_array":        (None,    "plain", "plain", False, True),
    "gs_init_array":          (None,    "plain", None,    False, True),
    "gs_gather":              ("local", "plain", "plain", True,  True),
    "gs_scatter":             ("local", "plain", "plain", True,  False),
    "gs_init":                ("local", "plain", None,    False, True),
    "gs_gather_vec":          ("local", "vec",   "vec",   True,  True),
    "gs_scatter_vec":         ("local", "vec",   "vec",   True,  False),
    "gs_init_vec":            ("local", "vec",   None,    False, True),
    "gs_gather_many":         ("local", "many",  "many",  True,  True),
    "gs_scatter_many":        ("local", "many",  "many",  True,  False),
    "gs_init_many":           ("local", "many",  None,    False, True),
    "gs_gather_vec_to_many":  ("recv",  "many",  "vec",   False, True),
    "gs_scatter_many_to_vec": ("send",  "vec",   "many",  False, False),
    "gs_scatter_vec_to_many": ("send",  "many",  "vec",   False, False),
}


def _field(layout, n, dt, op, seed):
    if layout == "many":
        return [cases.values(n, dt, op, seed + k) for k in range(VN)]
    return cases.values(n * (VN if layout == "vec" else 1), dt, op, seed)


def make(name, mp_table, dt, op, seed=5):
    """-> (out, inp, vn, map): numpy arrays (lists of arrays for "many"); inp is
    out itself where the reference's callers alias them (gs_aux: gather and
    scatter run in place on u, src/gs.upc:1181-1184)."""
    mname, lo, li, alias, _ = ENTRIES[name]
    if mname is None:
        out = cases.values(NARR, dt, op, seed)
        inp = cases.values(NARR, dt, op, seed + 1) if li else None
        return out, inp, NARR, None
    mp, n = mp_table[mname]
    first, rest = extents(mp)
    size = max(first, rest, n)
    gather = "gather" in name or "init" in name       # first entries index `out`
    n_out = size if alias else max(1, first if gather else rest)
    n_in = size if alias else max(1, rest if gather else first)
    out = _field(lo, n_out, dt, op, seed)
    inp = out if alias else (_field(li, n_in, dt, op, seed + 50) if li else None)
    return out, inp, VN, mp
