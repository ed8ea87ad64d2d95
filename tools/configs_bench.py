"""Device-resident timing of the BASELINE.json configurations beyond the headline one
(config 3: gs_vec / gs_many with 3 components in float and double; other domains/ops;
the CSR path used for flagged labels).  Prints one line per case with algorithmic bytes
and fraction of the measured HBM peak.   python tools/configs_bench.py [--ez 48]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "exaflow-exags-upc_b200")]
import torch  # noqa: E402
import exags_b200 as X  # noqa: E402
from exags_b200 import semmesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ez", type=int, default=48)
ap.add_argument("--order", type=int, default=7)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
peak = 6556.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dims = (64, 64, args.ez, args.order)
ids = semmesh.box_ids(*dims)
n = ids.size
g = X.gs_setup(ids)
info = g.info()
G, S = info["groups"], info["members"]
stream = torch.cuda.current_stream().cuda_stream
DT = {X.gs_double: torch.float64, X.gs_float: torch.float32, X.gs_int: torch.int32, X.gs_long: torch.int64}


def sectors(esz, vn):
    # distinct 32 B sectors touched: every sector at N=7 for 4- and 8-byte types (computed for f64 by the library)
    return info["sectors_f64"] * esz * vn / 8.0 if esz * vn >= 8 else info["sectors_f64"] / 2.0


def run(name, mode, vn, dom, op, handle=g, idx_bytes=None):
    dt = DT[dom]
    esz = torch.empty(0, dtype=dt).element_size()
    if mode == X.mode_many:
        u = [torch.ones(n, dtype=dt, device="cuda") for _ in range(vn)]
    else:
        u = torch.ones(n * vn, dtype=dt, device="cuda")
    call = lambda: X.gs_async(u, mode, vn, dom, op, 0, handle, stream)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    b_alg = 2 * 32 * sectors(esz, vn) + (4 * S + 4 * (G + 1) if idx_bytes is None else idx_bytes)
    print(f"{name:34s} {ms:8.4f} ms  {n / ms / 1e6:8.1f} GDOF/s  B_alg {b_alg / n:6.2f} B/DOF  "
          f"{b_alg / ms / 1e6:7.0f} GB/s  frac {b_alg / ms / 1e6 / peak:.3f}", flush=True)


print(f"mesh 64x64x{args.ez} N={args.order}: n={n} G={G} S={S}")
run("gs(add,double)", X.mode_plain, 1, X.gs_double, X.gs_add)
run("gs(add,float)", X.mode_plain, 1, X.gs_float, X.gs_add)
run("gs(max,int)", X.mode_plain, 1, X.gs_int, X.gs_max)
run("gs(min,long)", X.mode_plain, 1, X.gs_long, X.gs_min)
run("gs(mul,double)", X.mode_plain, 1, X.gs_double, X.gs_mul)
run("gs_vec(3,add,double)", X.mode_vec, 3, X.gs_double, X.gs_add)
run("gs_vec(3,add,float)", X.mode_vec, 3, X.gs_float, X.gs_add)
run("gs_many(3,add,double)", X.mode_many, 3, X.gs_double, X.gs_add)
run("gs_many(3,add,float)", X.mode_many, 3, X.gs_float, X.gs_add)
# comm_dot on two device vectors of n doubles: 16 B per element, HBM-bound
va, vb = torch.ones(n, dtype=torch.float64, device="cuda"), torch.full((n,), 0.5, dtype=torch.float64, device="cuda")
for _ in range(3):
    X.comm_dot(va, vb)
import time as _t
torch.cuda.synchronize(); t0 = _t.perf_counter()
for _ in range(args.steps):
    r = X.comm_dot(va, vb)
ms = 1e3 * (_t.perf_counter() - t0) / args.steps
assert r == 0.5 * n
print(f"{'comm_dot(device, n doubles)':34s} {ms:8.4f} ms  (blocking call incl. the scalar D2H)  "
      f"{16 * n / ms / 1e6:7.0f} GB/s  frac {16 * n / ms / 1e6 / peak:.3f}", flush=True)
del va, vb
X.gs_free(g)
# flagged labels -> CSR path with per-group unflagged counts
ids2 = ids.copy()
ids2[::7] = -ids2[::7]
g2 = X.gs_setup(ids2)
i2 = g2.info()
G, S = i2["groups"], i2["members"]
run("gs(add,double) flagged 1-in-7 (sliced + init list)", X.mode_plain, 1, X.gs_double, X.gs_add, g2, idx_bytes=4 * S + 8 * G + 4)
X.gs_free(g2)
