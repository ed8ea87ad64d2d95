"""Weighted gs (exags_gs_weighted, include/exags_cuda.h) against the unfused pair it
replaces -- gs(add) followed by a multiplication of the whole field with the weight
(Nek's dssum + col2) -- device-resident on the M7 mesh.
    python tools/weighted_bench.py [--ez 48] [--steps 20]"""
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
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
peak = 6556.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ids = semmesh.box_ids(64, 64, args.ez, 7)
n = ids.size
g = X.gs_setup(ids)
del ids
info = g.info()
G, S, nsec = info["groups"], info["members"], info["sectors_f64"]
stream = torch.cuda.current_stream().cuda_stream


def timed(call):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


print(f"mesh 64x64x{args.ez} N=7: n={n} G={G} S={S}")
for dom, dt, esz in ((X.gs_double, torch.float64, 8), (X.gs_float, torch.float32, 4)):
    u = torch.ones(n, dtype=dt, device="cuda")
    m = torch.ones(n, dtype=dt, device="cuda")
    X.gs(m, dom, X.gs_add, 0, g)
    w = 1.0 / m                                    # inverse multiplicity: 1 off the shared nodes
    del m
    field = 2 * 32 * nsec * esz / 8.0              # every sector of the field read and written once
    idx = 4 * S + 4 * (G + 1)
    t_gs = timed(lambda: X.gs_async(u, X.mode_plain, 1, dom, X.gs_add, 0, g, stream))
    u.fill_(1.0)
    t_w = timed(lambda: X.gs_weighted(u, w, dom, X.gs_add, g, stream=stream))
    u.fill_(1.0)

    def pair():
        X.gs_async(u, X.mode_plain, 1, dom, X.gs_add, 0, g, stream)
        u.mul_(w)
    t_pair = timed(pair)
    b_w = field + idx + esz * S                    # + one weight per member
    name = "double" if esz == 8 else "float"
    print(f"gs(add,{name})                      {t_gs:8.4f} ms  {(field + idx) / t_gs / 1e6:7.0f} GB/s  frac {(field + idx) / t_gs / 1e6 / peak:.3f}")
    print(f"gs_weighted(add,{name})             {t_w:8.4f} ms  {b_w / t_w / 1e6:7.0f} GB/s  frac {b_w / t_w / 1e6 / peak:.3f}"
          f"  (B_alg = gs + {esz} B x S weights)")
    print(f"gs(add,{name}) then u *= w (torch)  {t_pair:8.4f} ms  -> fused saves {t_pair - t_w:.4f} ms ({100 * (1 - t_w / t_pair):.0f} %)", flush=True)
    del u, w
X.gs_free(g)
