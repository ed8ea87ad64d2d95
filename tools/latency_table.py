"""Per-call latency of gs(add,double) on small meshes -- where a CPU wins and where Nek's coarse-grid
handles live.  N=7 boxes from 512 DOF to 4 M DOF per rank:

    python tools/latency_table.py                                   # 1 rank
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/latency_table.py

Columns (us per call): blocking gs() on a device vector; gs_async back to back on a stream (launch
bound); the same calls replayed from a CUDA graph (10 calls per graph; with np > 1 this needs the
push transport, whose exchange number lives on the device); the reference's own CPU code on the same
mesh (oracle/_ref as forked UPC threads, its own timing protocol src/gs.upc:1094-1111), 1 thread per
rank.  The CPU arm runs first, before this process touches CUDA (fork)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "exaflow-exags-upc_b200"), os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
from exags_b200 import semmesh  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
BOXES = [(1, 1, 1), (2, 2, 2), (4, 4, 4), (8, 8, 8), (16, 16, 8), (16, 16, 32)]
N = 7


def slab_ids(box, r, np_):
    ex, ey, ez = box
    return semmesh.box_ids(ex, ey, ez, N, ez0=r * ez, Ez_global=ez * np_)


cpu_us = {}
if rank == 0:
    try:
        import ref
        if not ref.available():
            raise RuntimeError("oracle/_ref not built")
        for box in BOXES:
            ids = [slab_ids(box, r, world) for r in range(world)]
            data = [[np.ones(a.size) for a in ids]]
            reps = 200 if ids[0].size < 1_000_000 else 20
            _, times = ref.run(ids, [(0, 1, 0, 0, 0)], data, method=1, reps=reps)
            cpu_us[box] = 1e6 * float(times[0])
    except Exception as exc:                                    # noqa: BLE001
        print(f"(reference CPU arm unavailable: {exc})")

import torch  # noqa: E402
import exags_b200 as X  # noqa: E402

torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X.lib().exags_set_device(local)


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = 1e6 * (time.perf_counter() - t0) / reps
    dev = 1e3 * e0.elapsed_time(e1) / reps
    t = torch.tensor([wall, dev], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), float(t[1])


if rank == 0:
    print(f"gs(add,double), N={N}, {world} rank(s), transport {X.exchange_transport()}; microseconds per call")
    print(f"{'elements/rank':>14s} {'DOF/rank':>10s} {'gs() blocking':>14s} {'gs_async stream':>16s} {'CUDA graph x10':>15s} {'reference CPU':>14s}")
for box in BOXES:
    ids = slab_ids(box, rank, world)
    g = X.gs_setup(ids)
    u = torch.ones(ids.size, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    X.gs(u, X.gs_double, X.gs_max, 0, g)
    reps = 200
    blk, _ = timed(lambda: X.gs(u, X.gs_double, X.gs_max, 0, g), reps)
    _, asy = timed(lambda: X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_max, 0, g, st), reps)
    gr = float("nan")
    if world == 1 or X.exchange_transport() == "push":
        side = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        barrier()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(10):
                X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_max, 0, g, side.cuda_stream)
        _, gr = timed(graph.replay, 20)
        gr /= 10
        del graph
    if rank == 0:
        c = cpu_us.get(box)
        print(f"{'x'.join(map(str, box)):>14s} {ids.size:10d} {blk:14.1f} {asy:16.1f} {gr:15.1f} "
              f"{(f'{c:14.1f}' if c is not None else '           n/a')}", flush=True)
    X.gs_free(g)
if dist is not None:
    dist.destroy_process_group()
