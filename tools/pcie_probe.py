"""How fast can every rank move 805 MB host->device and device->host at once, and both ways
at once?  (Explains the e2e numbers of bench.py at N>1: GPUs behind one PCIe switch share its uplink.)
    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py"""
import os
import time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 100663296
h1 = torch.empty(n, dtype=torch.float64).pin_memory(); h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.empty_like(d1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()


def run(what):
    sync(); t0 = time.perf_counter()
    for _ in range(3):
        if what in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
        if what in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    sync()
    return (time.perf_counter() - t0) / 3


for what in ("h2d", "d2h", "both"):
    run(what)
    t = run(what)
    if rank == 0:
        print(f"N={world} {what}: {t*1e3:.2f} ms per 805 MB each way -> {n*8/t/1e9:.1f} GB/s per GPU per direction", flush=True)
if world > 1:
    dist.destroy_process_group()
