"""A/B timing on M7 of the knobs that are read at gs_setup time, one handle per setting, interleaved
so that drift hits every setting alike:  python tools/knob_bench.py
    EXAGS_SELL_LINE_SHIFT=4|5|6   chain ordering on 16/32/64-element lines (default 5)
for plain gs on double / float / int, gs_vec(3) and gs_many(3) in float and double, the weighted gs, and
a unique=1 handle (both transposes)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "exaflow-exags-upc_b200")]
import torch  # noqa: E402
import exags_b200 as X  # noqa: E402
from exags_b200 import semmesh  # noqa: E402

peak = 6556.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ids = semmesh.box_ids(64, 64, 48, 7)
n = ids.size
KNOBS = ("EXAGS_SELL_LINE_SHIFT",)


def handle(unique=0, **env):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    g = X.gs_setup(ids, unique=unique)
    for k in KNOBS:
        os.environ.pop(k, None)
    return g


H = {"line_shift=5": handle(), "line_shift=4": handle(EXAGS_SELL_LINE_SHIFT=4), "line_shift=6": handle(EXAGS_SELL_LINE_SHIFT=6)}
info = H["line_shift=5"].info()
G, S, nsec = info["groups"], info["members"], info["sectors_f64"]
stream = torch.cuda.current_stream().cuda_stream
for label, mode, vn, dom, dt in (("gs(add,double)", X.mode_plain, 1, X.gs_double, torch.float64),
                                 ("gs(add,float)", X.mode_plain, 1, X.gs_float, torch.float32),
                                 ("gs(max,int)", X.mode_plain, 1, X.gs_int, torch.int32),
                                 ("gs_vec(3,add,float)", X.mode_vec, 3, X.gs_float, torch.float32),
                                 ("gs_vec(3,add,double)", X.mode_vec, 3, X.gs_double, torch.float64),
                                 ("gs_many(3,add,float)", X.mode_many, 3, X.gs_float, torch.float32),
                                 ("gs_many(3,add,double)", X.mode_many, 3, X.gs_double, torch.float64),
                                 ("gs_weighted(add,double)", "w", 1, X.gs_double, torch.float64),
                                 ("gs_weighted(add,float)", "w", 1, X.gs_float, torch.float32)):
    esz = torch.empty(0, dtype=dt).element_size()
    b_alg = 2 * 32 * (nsec * esz * vn / 8.0 if esz * vn >= 8 else nsec / 2.0) + 4 * S + 4 * (G + 1) + (esz * S if mode == "w" else 0)
    if mode == X.mode_many:
        u = [torch.ones(n, dtype=dt, device="cuda") for _ in range(vn)]
    else:
        u = torch.ones(n * vn, dtype=dt, device="cuda")
    wt = torch.full((n,), 0.5, dtype=dt, device="cuda") if mode == "w" else None
    op = X.gs_add if dt.is_floating_point else X.gs_max

    def call(g):
        if mode == "w":
            X.gs_weighted(u, wt, dom, op, g, stream)
        else:
            X.gs_async(u, mode, vn, dom, op, 0, g, stream)
    best = {}
    for rep in range(3):
        for name, g in H.items():
            for _ in range(3):
                call(g)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                call(g)
            e1.record()
            torch.cuda.synchronize()
            best.setdefault(name, []).append(e0.elapsed_time(e1) / 20)
            for x in (u if isinstance(u, list) else [u]):
                x.fill_(1)
    for name, ts in best.items():
        ms = sorted(ts)[len(ts) // 2]
        print(f"{label:24s} {name:22s} {ms:.4f} ms (median of {len(ts)}: {' '.join(f'{t:.4f}' for t in ts)})  "
              f"frac {b_alg / ms / 1e6 / peak:.3f}", flush=True)

for g in H.values():
    X.gs_free(g)
U = {"line_shift=5": handle(unique=1), "line_shift=4": handle(unique=1, EXAGS_SELL_LINE_SHIFT=4)}
u = torch.ones(n, dtype=torch.float64, device="cuda")
for t in (0, 1):
    best = {}
    for rep in range(3):
        for name, g in U.items():
            for _ in range(3):
                X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_add, t, g, stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_add, t, g, stream)
            e1.record()
            torch.cuda.synchronize()
            best.setdefault(name, []).append(e0.elapsed_time(e1) / 20)
            u.fill_(1)
    for name, ts in best.items():
        print(f"unique=1 gs(add,double) t={t}   {name:22s} {sorted(ts)[1]:.4f} ms ({' '.join(f'{x:.4f}' for x in ts)})", flush=True)
