"""gs_setup wall time with several ranks (z-slabs of an N=7 box), host builders
against the device passes.  Ranks are spawned here; with fewer GPUs than ranks
they share devices (the library then exchanges over CUDA IPC).
  python tools/setup_bench_mp.py --np 2 --ez 24        # 2 x 50 M DOF"""
import argparse
import os
import subprocess
import sys
import time
import uuid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "exaflow-exags-upc_b200"))


def worker(a):
    import torch
    import exags_b200 as X
    from exags_b200 import semmesh
    rank, np_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    X.lib().exags_set_device(rank % torch.cuda.device_count())
    ids = semmesh.box_ids(64, 64, a.ez, 7, ez0=a.ez * rank, Ez_global=a.ez * np_)
    X.comm_world()
    for mode in ("host", "device", "device"):
        os.environ["EXAGS_SETUP"] = mode
        X.lib().exags_comm_barrier(X.comm_world())
        t0 = time.perf_counter()
        g = X.gs_setup(ids)
        torch.cuda.synchronize()
        X.lib().exags_comm_barrier(X.comm_world())
        t1 = time.perf_counter()
        if rank == 0:
            i = g.info()
            print(f"np={np_} n/rank={ids.size} {mode:7s}: {t1 - t0:.3f} s  (boundary {i['boundary']}, "
                  f"interior groups {i['interior_groups']}, sell entries {i['sell_entries']})", flush=True)
        X.gs_free(g)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--np", type=int, default=2)
    ap.add_argument("--ez", type=int, default=24)
    a = ap.parse_args()
    if "RANK" in os.environ and os.environ.get("EXAGS_BENCH_CHILD") == "1":
        worker(a)
        sys.exit(0)
    key = uuid.uuid4().hex[:16]
    procs = []
    for r in range(a.np):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(a.np), LOCAL_RANK=str(r), EXAGS_JOB_KEY=key,
                   EXAGS_BENCH_CHILD="1", EXAGS_SETUP_PROFILE="1")
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--np", str(a.np), "--ez", str(a.ez)], env=env))
    sys.exit(max(p.wait() for p in procs))
