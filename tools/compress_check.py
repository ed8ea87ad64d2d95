"""EXAGS_SELL_COMPRESS=1 (16-bit index stream, csrc/sell_window.h) on the GPU: parity of the
compressed kernel against the raw one on a SEM box (every dom x op, bit for bit), then the M7
timing of both.  Written in round 1 without GPU time left: run this first in round 2
(`python tools/compress_check.py`), then add the parity part to tests/test_gs_gpu.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "exaflow-exags-upc_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import exags_b200 as X  # noqa: E402
from exags_b200 import semmesh  # noqa: E402
import cases  # noqa: E402


def handles(ids):
    os.environ["EXAGS_SELL_COMPRESS"] = "0"
    raw = X.gs_setup(ids)
    os.environ["EXAGS_SELL_COMPRESS"] = "1"
    comp = X.gs_setup(ids)
    os.environ["EXAGS_SELL_COMPRESS"] = "0"
    return raw, comp


bad = 0
for dims in ((6, 5, 4, 7), (12, 12, 8, 3), (5, 4, 3, 9)):
    ids = semmesh.box_ids(*dims)
    perm = np.random.default_rng(1).permutation(dims[0] * dims[1] * dims[2])
    for name, lab in (("box", ids), ("box, elements in random order", ids.reshape(perm.size, -1)[perm].reshape(-1))):
        raw, comp = handles(lab)
        for dname, dt in cases.DTYPES.items():
            for op in cases.OPS:
                u = cases.values(lab.size, dt, op, 5)
                a, b = torch.from_numpy(u.copy()).cuda(), torch.from_numpy(u.copy()).cuda()
                opi, dom = cases.OPS.index(op), X.NP_DOM[np.dtype(dt)]
                X.gs(a, dom, opi, 0, raw)
                X.gs(b, dom, opi, 0, comp)
                if not torch.equal(a, b):
                    bad += 1
                    print(f"MISMATCH {dims} {name} {dname} {op}")
        X.gs_free(raw); X.gs_free(comp)
print("parity:", "ok" if not bad else f"{bad} mismatches", flush=True)

peak = 6556.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ids = semmesh.box_ids(64, 64, 48, 7)
n = ids.size
raw, comp = handles(ids)
del ids
info = raw.info()
b_alg = 2 * 32 * info["sectors_f64"] + 4 * info["members"] + 4 * (info["groups"] + 1)
stream = torch.cuda.current_stream().cuda_stream
u = torch.ones(n, dtype=torch.float64, device="cuda")
for name, g in (("raw 32-bit indices", raw), ("16-bit codes", comp), ("raw 32-bit indices", raw), ("16-bit codes", comp)):
    for _ in range(3):
        X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_add, 0, g, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_add, 0, g, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    u.fill_(1.0)
    print(f"gs(add,double) M7, {name:20s} {ms:.4f} ms  {b_alg / ms / 1e6:6.0f} GB/s of B_alg  frac {b_alg / ms / 1e6 / peak:.3f}", flush=True)
sys.exit(1 if bad else 0)
