"""One gs(add,double) call of the weak-scaling configuration (M7 slab per rank) between
cudaProfilerStart/Stop, for an ncu capture of the np > 1 kernels with NVLink counters.
Launched by tools/ncu_np2.sh: rank 0 under ncu, the other ranks plain."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "exaflow-exags-upc_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import exags_b200 as X  # noqa: E402
from exags_b200 import semmesh  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)
X.lib().exags_set_device(rank)
EZ = int(os.environ.get("NP2_EZ", "48"))
ids = semmesh.box_ids(64, 64, EZ, 7, ez0=rank * EZ, Ez_global=EZ * world)
g = X.gs_setup(ids)
info = g.info()
u = torch.ones(ids.size, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_add, 0, g, st)
torch.cuda.synchronize()
dist.barrier()
torch.cuda.profiler.start()
X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_add, 0, g, st)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
dist.barrier()
if rank == 0:
    print(f"np={world} transport={X.exchange_transport()} send_slots={info['send_slots']} "
          f"recv_slots={info['recv_slots']} boundary_groups={info['boundary']}", flush=True)
X.gs_free(g)
dist.destroy_process_group()
