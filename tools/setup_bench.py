"""gs_setup wall time, host builders (EXAGS_SETUP=host) against the device
passes, on the M7 mesh (or --dims ex ey ez N).  Prints the library's own stage
profile (EXAGS_SETUP_PROFILE=1) and the wall time of each call."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "exaflow-exags-upc_b200"))

import numpy as np
import torch

import exags_b200 as X
from exags_b200 import semmesh

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs=4, default=[64, 64, 48, 7])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
os.environ["EXAGS_SETUP_PROFILE"] = "1"
ids = semmesh.box_ids(*a.dims)
print(f"n = {ids.size}", flush=True)
torch.cuda.init()
X.comm_world()
for mode in ("host", "device", "device-resident labels"):
    os.environ["EXAGS_SETUP"] = "host" if mode == "host" else "device"
    src = torch.from_numpy(ids).cuda() if mode.startswith("device-") else ids
    for r in range(a.reps if mode != "host" else 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = X.gs_setup(src)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        print(f"{mode:24s} rep {r}: {t1 - t0:.3f} s   {g.info()['groups']} groups", flush=True)
        X.gs_free(g)
