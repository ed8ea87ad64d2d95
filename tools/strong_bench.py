#!/usr/bin/env python
"""BASELINE.json config 5: strong scaling of one fixed N=9 box mesh over P GPUs
with an irregular ("crystal-router style") block partition, every gs_dom x
gs_op.  Elements are grouped in bx x by x bz blocks and block b goes to rank
perm[b] % P (fixed LCG shuffle), so every rank neighbours every
other and message sizes are ragged.

    python -m torch.distributed.run --nproc-per-node P tools/strong_bench.py [--box 100 100 80]

Each case is checked, at full size, against an independent computation that
does not depend on the partition: a dense scatter-reduce by global id on the
GPU (torch), combined over ranks with torch.distributed, gathered back by id.
Integer results and min/max must match exactly; floating add/mul are compared
with a relative bound because the dense sum associates differently (bit parity
with the reference's order is what tests/ prove against the oracle).

Prints one JSON line per case on rank 0, then a summary line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "exaflow-exags-upc_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import exags_b200 as X  # noqa: E402
from exags_b200 import semmesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--box", type=int, nargs=3, default=[100, 100, 80])
ap.add_argument("--block", type=int, nargs=3, default=[10, 10, 8])
ap.add_argument("--order", type=int, default=9)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-check", action="store_true")
ap.add_argument("--out", default="")
ap.add_argument("--only", default="", help="run just this case, e.g. add,double")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X.lib().exags_set_device(local)
Ex, Ey, Ez = args.box
N = args.order
m3 = (N + 1) ** 3

t0 = time.perf_counter()
ids, elems = semmesh.block_partition_ids(Ex, Ey, Ez, N, *args.block, world, rank)
n = ids.size
t1 = time.perf_counter()
g = X.gs_setup(ids, method=X.gs_pairwise)
t_setup = time.perf_counter() - t1
info = g.info()
dof = (elems.astype(np.uint64)[:, None] * np.uint64(m3) + np.arange(m3, dtype=np.uint64)[None, :]).reshape(-1)
# one partition-independent 53-bit hash per DOF (of its global DOF number); every field derives from it on the GPU
with np.errstate(over="ignore"):
    hbits = semmesh._splitmix64((dof * np.uint64(0x9E3779B97F4A7C15)) ^ np.uint64(semmesh.SEED)) >> np.uint64(11)
del dof
hdev = torch.from_numpy(hbits.astype(np.int64)).cuda()
del hbits


def make_field(tdt, opname):
    """FP: uniform [0.5,1.5); int add/min/max: [-1000,1000]; int mul: {1,2}"""
    if tdt.is_floating_point:
        return (0.5 + hdev.to(torch.float64) * (1.0 / (1 << 53))).to(tdt)
    if opname == "mul":
        return (1 + (hdev & 1)).to(tdt)
    return (hdev % 2001 - 1000).to(tdt)
ids_dev = torch.from_numpy(ids).cuda() - 1           # dense slot of each local DOF
nglob = (Ex * N + 1) * (Ey * N + 1) * (Ez * N + 1)
del ids
stream = torch.cuda.current_stream().cuda_stream
n_total = Ex * Ey * Ez * m3

DT = {X.gs_double: (torch.float64, np.float64), X.gs_float: (torch.float32, np.float32),
      X.gs_int: (torch.int32, np.int32), X.gs_long: (torch.int64, np.int64)}
NAMES = {X.gs_double: "double", X.gs_float: "float", X.gs_int: "int", X.gs_long: "long"}
OPS = {X.gs_add: ("add", "sum"), X.gs_mul: ("mul", "prod"), X.gs_min: ("min", "amin"), X.gs_max: ("max", "amax")}
RED = {"sum": "SUM", "prod": "PRODUCT", "amin": "MIN", "amax": "MAX"}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def expected(u, opname):
    """dense scatter-reduce by global id, all ranks combined, gathered back"""
    red = OPS[opname][1]
    init = {"sum": 0, "prod": 1}.get(red)
    if init is None:
        lim = torch.finfo(u.dtype) if u.dtype.is_floating_point else torch.iinfo(u.dtype)
        init = lim.max if red == "amin" else lim.min
    dense = torch.full((nglob,), init, dtype=u.dtype, device="cuda")
    dense.scatter_reduce_(0, ids_dev, u, red, include_self=True)
    if world > 1:
        dist.all_reduce(dense, op=getattr(dist.ReduceOp, RED[red]))
    return dense[ids_dev]


results = []
ok_all = True
for dom in (X.gs_double, X.gs_float, X.gs_int, X.gs_long):
    tdt, ndt = DT[dom]
    for op in (X.gs_add, X.gs_mul, X.gs_min, X.gs_max):
        opname = OPS[op][0]
        if args.only and args.only != f"{opname},{NAMES[dom]}":
            continue
        u0 = make_field(tdt, opname)
        u = u0.clone()
        X.gs_async(u, X.mode_plain, 1, dom, op, 0, g, stream)
        torch.cuda.synchronize()
        status = "unchecked"
        if not args.no_check:
            want = expected(u0, op)
            if tdt.is_floating_point and opname in ("add", "mul"):
                rel = ((u - want).abs() / want.abs()).max().item()
                eps = torch.finfo(tdt).eps
                good = rel <= 8 * eps
                status = f"max rel diff vs dense reduce {rel / eps:.2f} eps"
            else:
                bad = int((u != want).sum().item())
                good = bad == 0
                status = "exact" if good else f"{bad} DOF differ"
            if world > 1:
                f = torch.tensor([0 if good else 1], device="cuda")
                dist.all_reduce(f)
                good = int(f.item()) == 0
            ok_all &= good
            del want
        # timing: the field is far larger than L2, calls back to back
        for _ in range(3):
            X.gs_async(u, X.mode_plain, 1, dom, op, 0, g, stream)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            X.gs_async(u, X.mode_plain, 1, dom, op, 0, g, stream)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        line = {"case": f"gs({opname},{NAMES[dom]})", "n_gpus": world, "ms": ms,
                "GDOF/s": n_total / ms / 1e6, "check": status}
        results.append(line)
        if rank == 0:
            print(json.dumps(line), flush=True)
        del u, u0

if rank == 0:
    summ = {"config": f"{Ex}x{Ey}x{Ez} hex N={N} ({n_total} DOF total), blocks {args.block}, P={world}",
            "local_dof_rank0": n, "setup_s": t_setup, "neighbours_rank0": info["neighbours"],
            "send_slots_rank0": info["send_slots"], "boundary_groups_rank0": info["boundary"],
            "interior_groups_rank0": info["interior_groups"], "all_checks_ok": bool(ok_all),
            "exchange": os.environ.get("EXAGS_EXCHANGE", "default"), "cases": results}
    print(json.dumps(summ), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(summ, fh, indent=1)
X.gs_free(g)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
