"""BASELINE config 3 at full size (M7, 3 components): gs_vec and gs_many must give, component
by component, exactly what three plain gs calls give (same association order), in float and
double; plus gs_weighted against gs followed by the multiplication.  A size-independent
property check for sizes the CPU oracle cannot reach; run it on the GPU box
(`python tools/fullsize_vec_check.py`), then move it into tests/test_gs_gpu.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "exaflow-exags-upc_b200")]
import torch  # noqa: E402
import exags_b200 as X  # noqa: E402
from exags_b200 import semmesh  # noqa: E402

ids = semmesh.box_ids(64, 64, 48, 7)
n = ids.size
g = X.gs_setup(ids)
del ids
bad = 0
for dom, dt in ((X.gs_double, torch.float64), (X.gs_float, torch.float32)):
    gen = torch.Generator(device="cuda").manual_seed(7)
    comps = [torch.rand(n, dtype=dt, device="cuda", generator=gen) + 0.5 for _ in range(3)]
    want = [c.clone() for c in comps]
    for c in want:
        X.gs(c, dom, X.gs_add, 0, g)
    many = [c.clone() for c in comps]
    X.gs_many(many, 3, dom, X.gs_add, 0, g)
    ok_many = all(torch.equal(a, b) for a, b in zip(many, want))
    del many
    vec = torch.stack(comps, dim=1).contiguous().view(-1)
    X.gs_vec(vec, 3, dom, X.gs_add, 0, g)
    ok_vec = torch.equal(vec.view(n, 3), torch.stack(want, dim=1))
    del vec
    w = torch.rand(n, dtype=dt, device="cuda", generator=gen) * 0.75 + 0.25
    m = torch.ones(n, dtype=dt, device="cuda")
    X.gs(m, dom, X.gs_add, 0, g)
    shared = m > 1
    del m
    fused = comps[0].clone()
    X.gs_weighted(fused, w, dom, X.gs_add, g)
    ok_w = torch.equal(fused, torch.where(shared, want[0] * w, want[0]))
    print(f"{dt}: gs_many == 3 x gs: {ok_many}; gs_vec == 3 x gs: {ok_vec}; gs_weighted == gs then *w: {ok_w}", flush=True)
    bad += (not ok_many) + (not ok_vec) + (not ok_w)
    del comps, want, fused, w, shared
X.gs_free(g)
sys.exit(1 if bad else 0)
