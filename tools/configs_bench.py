"""Device-resident timing of the BASELINE.json configurations beyond the headline one
(config 3: gs_vec / gs_many with 3 components in float and double; other domains/ops; maps with
flagged labels; unique=1 handles as amg.upc builds them; the weighted gs; comm_dot) on M7.
One line per case: ms, GDOF/s, algorithmic bytes and fraction of the measured HBM peak.

    python tools/configs_bench.py [--ez 48] [--steps 20] [--only SUBSTR] [--ncu]

--ncu: every case is warmed up as usual and then run ONCE between cudaProfilerStart/Stop, so
    ncu --profile-from-start off --set full ... python tools/configs_bench.py --ncu
captures exactly one launch of every kernel on the list (no timing is printed in that mode).

Algorithmic bytes: 32 B for every distinct sector of the field(s)
holding a member that is read, 32 B for every distinct sector holding a member that is written,
4 B per member index and per group offset, plus what the variant adds (one weight per member;
one identity store per flagged single).  For maps with flagged labels the read and written member
sets differ (src/gs.upc:1181-1184), so both are counted from the handle's own maps."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "exaflow-exags-upc_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import exags_b200 as X  # noqa: E402
from exags_b200 import semmesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ez", type=int, default=48)
ap.add_argument("--order", type=int, default=7)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--only", default="")
ap.add_argument("--ncu", action="store_true")
args = ap.parse_args()
peak = 6556.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dims = (64, 64, args.ez, args.order)
ids = semmesh.box_ids(*dims)
n = ids.size
stream = torch.cuda.current_stream().cuda_stream
DT = {X.gs_double: torch.float64, X.gs_float: torch.float32, X.gs_int: torch.int32, X.gs_long: torch.int64}
NIL = 0xFFFFFFFF


def sector_counts(handle, esz, vn_tuple, transpose):
    """(sectors read, sectors written, members, groups, flagged singles) of one call on a T[n][vn_tuple]
    field, from the handle's own sentinel maps: map 1 = every group in full, map 0 = unflagged members only.
    transpose 0 reads map 0's members and writes map 1's; transpose 1 the other way round."""
    def members_of(which, heads=False):
        m = torch.from_numpy(handle.dump_map(which).astype(np.int64)).cuda()
        if heads:                                     # first entry of every group: the primaries
            prev = torch.cat([torch.tensor([NIL], device="cuda"), m[:-1]])
            return m[(prev == NIL) & (m != NIL)]
        return m[m != NIL], int((m == NIL).sum().item()) - 1

    def nsec(idx):
        if idx.numel() == 0:
            return 0
        lo = (idx * (esz * vn_tuple)) // 32
        hi = (idx * (esz * vn_tuple) + esz * vn_tuple - 1) // 32
        mark = torch.zeros(int(hi.max().item()) + 1, dtype=torch.bool, device="cuda")
        for k in range(int((hi - lo).max().item()) + 1):
            mark[torch.minimum(lo + k, hi)] = True
        return int(mark.sum().item())
    full, G = members_of(1)
    info = handle.info()
    unf, _ = members_of(0)
    if unf.numel() == full.numel() and info["flagged_primaries"] == 0:
        s = nsec(full)
        return s, s, full.numel(), G, 0
    fp = torch.from_numpy(handle.dump_map(2).astype(np.int64)).cuda()
    fp = fp[fp != NIL]
    # a store into a sector makes DRAM read it first (32-byte ECC granularity) unless the whole sector is
    # written, so the sectors read are those of every member touched at all; written: those of the stored ones
    # transpose 1: the gather over map 1 writes every primary, the scatter over map 0 the unflagged members
    touched = torch.cat([full, fp])
    wr = touched if not transpose else torch.cat([members_of(1, heads=True), unf])
    return nsec(touched), nsec(wr), full.numel(), G, info["flagged_singles"]


def run(name, mode, vn, dom, op, handle, transpose=0, weighted=False, secs=None):
    if args.only and args.only not in name:
        return
    dt = DT[dom]
    esz = torch.empty(0, dtype=dt).element_size()
    if mode == X.mode_many:
        u = [torch.ones(n, dtype=dt, device="cuda") for _ in range(vn)]
    else:
        u = torch.ones(n * vn, dtype=dt, device="cuda")
    w = torch.full((n,), 0.5, dtype=dt, device="cuda") if weighted else None
    if weighted and mode == X.mode_vec:
        call = lambda: X.gs_weighted_vec(u, vn, w, dom, op, handle, stream)
    elif weighted and mode == X.mode_many:
        call = lambda: X.gs_weighted_many(u, vn, w, dom, op, handle, stream)
    elif weighted:
        call = lambda: X.gs_weighted(u, w, dom, op, handle, stream)
    else:
        call = lambda: X.gs_async(u, mode, vn, dom, op, transpose, handle, stream)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    if args.ncu:
        torch.cuda.profiler.start()
        call()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(f"{name}: captured", flush=True)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    tuple_ = vn if mode == X.mode_vec else 1
    comps = vn if mode == X.mode_many else 1
    key = (id(handle), esz, tuple_, transpose)
    if key not in SEC:
        SEC[key] = sector_counts(handle, esz, tuple_, transpose)
    rd, wr, S, G, nfs = SEC[key]
    extra = (esz * S if weighted else 0) + 4 * nfs
    b_sec = 32 * (rd + wr) * comps + 4 * S + 4 * (G + 1) + extra
    # SURVEY.md 8d's per-unit figure ("every sector is touched": 2*vn*sizeof(T)*n + 4S + 4(G+1), with the
    # library's own count of touched sectors of a double field scaled to the element size); for gs_vec tuples
    # that over-counts -- 43 % of the DOFs are no members and a 24-byte tuple leaves whole sectors untouched --
    # so the count from the map is printed beside it
    sec_f64 = handle.info()["sectors_f64"]
    vtot = vn if mode != X.mode_plain else 1
    b_svy = 2 * 32 * (sec_f64 * esz * vtot / 8.0 if esz * vtot >= 8 else sec_f64 / 2.0) + 4 * S + 4 * (G + 1) + extra
    sym = rd == wr
    print(f"{name:42s} {ms:8.4f} ms {n / ms / 1e6:7.1f} GDOF/s | SURVEY 8d bytes {b_svy / n:6.2f} B/DOF frac "
          f"{(b_svy / ms / 1e6 / peak if sym else float('nan')):.3f} | sectors from the map {b_sec / n:6.2f} B/DOF "
          f"frac {b_sec / ms / 1e6 / peak:.3f}", flush=True)


SEC = {}
g = X.gs_setup(ids)
info = g.info()
print(f"mesh 64x64x{args.ez} N={args.order}: n={n} G={info['groups']} S={info['members']}")
run("gs(add,double)", X.mode_plain, 1, X.gs_double, X.gs_add, g)
run("gs(add,float)", X.mode_plain, 1, X.gs_float, X.gs_add, g)
run("gs(max,int)", X.mode_plain, 1, X.gs_int, X.gs_max, g)
run("gs(min,long)", X.mode_plain, 1, X.gs_long, X.gs_min, g)
run("gs(mul,double)", X.mode_plain, 1, X.gs_double, X.gs_mul, g)
run("gs_vec(3,add,double)", X.mode_vec, 3, X.gs_double, X.gs_add, g)
run("gs_vec(3,add,float)", X.mode_vec, 3, X.gs_float, X.gs_add, g)
run("gs_many(3,add,double)", X.mode_many, 3, X.gs_double, X.gs_add, g)
run("gs_many(3,add,float)", X.mode_many, 3, X.gs_float, X.gs_add, g)
run("gs_weighted(add,double)", X.mode_plain, 1, X.gs_double, X.gs_add, g, weighted=True)
run("gs_weighted(add,float)", X.mode_plain, 1, X.gs_float, X.gs_add, g, weighted=True)
run("gs_weighted_vec(3,add,double)", X.mode_vec, 3, X.gs_double, X.gs_add, g, weighted=True)
run("gs_weighted_many(3,add,double)", X.mode_many, 3, X.gs_double, X.gs_add, g, weighted=True)
run("gs_weighted_many(3,add,float)", X.mode_many, 3, X.gs_float, X.gs_add, g, weighted=True)
# comm_dot on two device vectors of n doubles: 16 B per element, HBM-bound
if not args.only or "comm_dot" in args.only:
    va, vb = torch.ones(n, dtype=torch.float64, device="cuda"), torch.full((n,), 0.5, dtype=torch.float64, device="cuda")
    for _ in range(3):
        X.comm_dot(va, vb)
    if args.ncu:
        torch.cuda.profiler.start(); X.comm_dot(va, vb); torch.cuda.profiler.stop()
    else:
        import time as _t
        torch.cuda.synchronize(); t0 = _t.perf_counter()
        for _ in range(args.steps):
            r = X.comm_dot(va, vb)
        ms = 1e3 * (_t.perf_counter() - t0) / args.steps
        assert r == 0.5 * n
        print(f"{'comm_dot(device, n doubles), blocking':46s} {ms:8.4f} ms  (incl. the scalar D2H)  "
              f"{16 * n / ms / 1e6:7.0f} GB/s  frac {16 * n / ms / 1e6 / peak:.3f}", flush=True)
    del va, vb
X.gs_free(g)
# the flagged case real callers produce: a unique=1 handle (src/amg.upc: gs_setup(..., 1, gs_auto, 1)).
# Every label keeps one unflagged copy: transpose 0 copies it to the others, transpose 1 sums into it.
if not args.only or "unique" in args.only:
    gu = X.gs_setup(ids, unique=1)
    run("unique=1 handle: gs(add,double) t=0", X.mode_plain, 1, X.gs_double, X.gs_add, gu, transpose=0)
    run("unique=1 handle: gs(add,double) t=1", X.mode_plain, 1, X.gs_double, X.gs_add, gu, transpose=1)
    run("unique=1 handle: gs(add,float) t=1", X.mode_plain, 1, X.gs_float, X.gs_add, gu, transpose=1)
    X.gs_free(gu)
# a synthetic pattern with flagged singles and partly flagged groups
if not args.only or "1-in-7" in args.only:
    ids2 = ids.copy()
    ids2[::7] = -ids2[::7]
    g2 = X.gs_setup(ids2)
    del ids2
    run("flagged 1-in-7: gs(add,double) t=0", X.mode_plain, 1, X.gs_double, X.gs_add, g2, transpose=0)
    run("flagged 1-in-7: gs(add,double) t=1", X.mode_plain, 1, X.gs_double, X.gs_add, g2, transpose=1)
    X.gs_free(g2)
