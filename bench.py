#!/usr/bin/env python
"""bench.py -- gs_op(add,double) throughput on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one gs(add,double) over one rank-local field of the synthetic
Nek5000-style mesh: at N=1 the M7 box 64x64x48, order 7
(100 663 296 DOF); at N>1 the same box per GPU as z-slabs of a 64x64x(48N)
mesh (weak scaling; pairwise halo exchange by peer stores over NVLink from the
boundary pack kernel, or NCCL send/recv with EXAGS_EXCHANGE=nccl).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput,
`e2e` the same call made with pinned HOST buffers through the public gs()
(H2D + kernels + D2H inside the timed region), `roofline` the fused kernel's
algorithmic bytes / CUDA-event time against the measured HBM peak, and
`cpu_baseline` the CPU oracle timed on this box's host cores on a bounded
sample.  --impl reference times the reference's CPU algorithm only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "exaflow-exags-upc_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

EX, EY, EZ, ORDER = 64, 64, 48, 7          # M7 per GPU
METRIC = "gs_op(add,double) throughput"
UNIT = "GDOF/s"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6
                          for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(sm)}


def nvlink_kib(index: int):
    """(tx, rx) payload KiB this GPU has moved over all its NVLink links so far -- the
    driver's hardware counters (NVML field values NVLINK_THROUGHPUT_DATA_TX/RX, scope = every
    link), or None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        # CUDA_VISIBLE_DEVICES may renumber: go through the UUID of the torch device
        import torch
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        h = None
        for k in range(pynvml.nvmlDeviceGetCount()):
            hk = pynvml.nvmlDeviceGetHandleByIndex(k)
            u = pynvml.nvmlDeviceGetUUID(hk)
            u = u.decode() if isinstance(u, bytes) else u
            if uuid in u:
                h = hk
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        all_links = 0xFFFFFFFF
        v = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, all_links),
                                                 (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, all_links)])
        if any(x.nvmlReturn != 0 for x in v):
            return None
        return int(v[0].value.ullVal), int(v[1].value.ullVal)
    except Exception:
        return None


# --------------------------------------------------------------------------
#  CPU arm: the oracle (port of the reference algorithm) on the host cores
# --------------------------------------------------------------------------
def _host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_arm(steps: int, warmup: int, seconds_budget: float = 25.0) -> dict:
    """Times the reference CPU path on a bounded sample of the workload: z-slabs
    of the same 64x64 N=7 mesh, one slab per rank, pairwise exchange, ranks on
    separate host cores.  Uses the reference's own code (oracle/_ref, built from
    /root/reference/src by oracle/build_ref.sh; ranks are forked UPC threads and
    timing follows src/gs.upc:1094-1111: barrier, 2 warm-ups, N timed calls,
    mean, max over ranks) when that library is present, else the oracle port
    with ranks on OpenMP threads."""
    from exags_b200 import semmesh
    cores = _host_cores()
    ranks = 1
    while ranks * 2 <= min(cores, 32):      # the reference runs gs on powers of two
        ranks *= 2
    ez_per = 2
    ids = [semmesh.box_ids(EX, EY, ez_per, ORDER, ez0=r * ez_per, Ez_global=ez_per * ranks)
           for r in range(ranks)]
    n_total = sum(a.size for a in ids)
    sample = (f"{EX}x{EY}x{ez_per * ranks} el N={ORDER} ({n_total} DOF) as {ranks} z-slab ranks, pairwise")
    try:
        import ref
        if not ref.available():
            raise RuntimeError("oracle/_ref not built")
        reps = max(3, min(steps, 50))          # as many timed calls as the GPU arm's steps
        data = [[semmesh.field_values(semmesh.slab_dof_numbers(EX, EY, ez_per, ORDER, r * ez_per), np.float64)
                 for r in range(ranks)]]
        # the shim's bump arena never frees and comm_alloc_thrd_buf keeps np mailboxes of the whole
        # receive volume per rank (src/comm.upc:174-179): size it by ranks^2; pages are only
        # committed when touched (MAP_NORESERVE)
        arena = 2 * ((1 << 30) + n_total * 100 + ranks * ranks * (8 << 20))
        try:
            memtotal = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
            arena = min(arena, memtotal // 2)
        except (ValueError, OSError):
            pass
        # the forked ranks share the box with whatever else runs on it: three runs of the reference's
        # protocol, the median reported (round 1's single run spread 2.7-4.0 GDOF/s between boxes)
        runs = []
        for _ in range(3):
            _, times = ref.run(ids, [(0, 1, 0, 0, 0)], data, method=1, reps=reps, arena_bytes=arena)
            runs.append(1e3 * float(times[0]))
        ms = sorted(runs)[1]
        return {"value": n_total / (ms * 1e-3) / 1e9, "unit": UNIT, "cores": ranks, "kind": "reference",
                "sample": sample + f", reference's own code as forked UPC threads, 2 warm + {reps} timed calls, "
                          f"median of 3 runs ({', '.join(f'{x:.1f}' for x in runs)} ms)",
                "ms_per_step": ms, "steps": reps}
    except Exception as exc:                    # fall back to the port
        note = f" (oracle/_ref unavailable: {exc})"
    import oracle as O
    os.environ["OMP_NUM_THREADS"] = str(ranks)
    orc = O.Oracle(ids, method=O.METHOD_PAIRWISE)
    us = [semmesh.field_values(semmesh.slab_dof_numbers(EX, EY, ez_per, ORDER, r * ez_per), np.float64)
          for r in range(ranks)]
    for _ in range(max(1, warmup)):
        orc.exec(us, O.M_PLAIN, 1, O.D_DOUBLE, O.O_ADD, 0)
    t_all, done = 0.0, 0
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.exec(us, O.M_PLAIN, 1, O.D_DOUBLE, O.O_ADD, 0)
        t_all += time.perf_counter() - t0
        done += 1
        if t_all > seconds_budget:
            break
    ms = 1e3 * t_all / done
    return {"value": n_total / (ms * 1e-3) / 1e9, "unit": UNIT, "cores": ranks, "kind": "port",
            "sample": sample + f", oracle port on OpenMP threads, {done} steps" + note,
            "ms_per_step": ms, "steps": done}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_arm(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": cb["steps"], "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"M7 {EX}x{EY}x{EZ} hex N={ORDER} per GPU, gs(add,double); "
                                   "each step a bounded sample: " + cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
#  GPU arm
# --------------------------------------------------------------------------
def run_check(X, g, world, n, ids_dev, ints_dev, hbuf, stream, dist) -> dict:
    """Exact, size-independent checks on the handle that was just timed, at every
    N (the reference runs its gs test at every thread count,
    tests/gs/gs_test.upc:118-240): none of them needs the CPU oracle.
      multiplicity   gs(add,double) of ones gives m in {1,2,4,8} and sum(1/m)
                     over all ranks is the number of distinct mesh nodes;
      add_long       gs(add,long) of a partition-independent random field
                     equals a dense scatter-add by global id (torch), summed
                     over ranks, gathered back by id -- bit for bit;
      max_idempotent gs(max,double) twice equals once;
      host_pointer   the multiplicity check through the blocking gs() on host
                     memory (the e2e path)."""
    import torch
    from exags_b200 import semmesh
    out = {}

    def allsum(t):
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    unique_nodes = semmesh.box_counts(EX, EY, EZ * world, ORDER)["unique"]
    m = torch.ones(n, dtype=torch.float64, device="cuda")
    X.gs_async(m, X.mode_plain, 1, X.gs_double, X.gs_add, 0, g, stream)
    torch.cuda.synchronize()
    in_set = bool(((m == 1) | (m == 2) | (m == 4) | (m == 8)).all().item())
    total = allsum((1.0 / m).sum().reshape(1))          # eighths: exact in double in any order
    out["multiplicity"] = in_set and float(total.item()) == float(unique_nodes)
    del m

    nglob = (EX * ORDER + 1) * (EY * ORDER + 1) * (EZ * world * ORDER + 1)
    dense = torch.zeros(nglob, dtype=torch.int64, device="cuda")
    dense.scatter_add_(0, ids_dev - 1, ints_dev)
    allsum(dense)
    want = dense[ids_dev - 1]
    del dense
    got = ints_dev.clone()
    X.gs_async(got, X.mode_plain, 1, X.gs_long, X.gs_add, 0, g, stream)
    torch.cuda.synchronize()
    out["add_long"] = bool(torch.equal(got, want))
    del got, want

    a = hbuf.cuda()
    X.gs_async(a, X.mode_plain, 1, X.gs_double, X.gs_max, 0, g, stream)
    torch.cuda.synchronize()
    b = a.clone()
    X.gs_async(b, X.mode_plain, 1, X.gs_double, X.gs_max, 0, g, stream)
    torch.cuda.synchronize()
    out["max_idempotent"] = bool(torch.equal(a, b))
    del a, b

    hbuf.fill_(1.0)
    X.gs(hbuf.numpy(), X.gs_double, X.gs_add, 0, g)
    in_set = bool(((hbuf == 1) | (hbuf == 2) | (hbuf == 4) | (hbuf == 8)).all().item())
    total = allsum((1.0 / hbuf).sum().reshape(1).cuda())
    out["host_pointer"] = in_set and float(total.item()) == float(unique_nodes)

    flags = torch.tensor([int(bool(v)) for v in out.values()], device="cuda")
    if dist is not None:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)     # a case passes only if it passed on every rank
    out = {k: bool(f) for k, f in zip(out, flags.tolist())}
    return {"ok": all(out.values()), "cases": out, "unique_nodes": unique_nodes,
            "what": "exact checks on the timed handle, all ranks (see run_check in bench.py)"}


def run_gpu(args) -> None:
    import torch
    import exags_b200 as X
    from exags_b200 import semmesh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
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

    # ---- mesh + setup ------------------------------------------------
    t0 = time.perf_counter()
    ids = semmesh.box_ids(EX, EY, EZ, ORDER, ez0=rank * EZ, Ez_global=EZ * world)
    n = ids.size
    t1 = time.perf_counter()
    g = X.gs_setup(ids, method=X.gs_pairwise, verbose=0)
    t_setup = time.perf_counter() - t1
    info = g.info()
    transport = X.exchange_transport()
    ids_dev = torch.from_numpy(ids).cuda()
    del ids
    dof = semmesh.slab_dof_numbers(EX, EY, EZ, ORDER, rank * EZ)
    host = torch.from_numpy(semmesh.field_values(dof, np.float64)).pin_memory()
    ints_dev = torch.from_numpy(semmesh.field_values(dof, np.int64)).cuda()
    del dof
    u = host.cuda()
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        X.gs_async(u, X.mode_plain, 1, X.gs_double, X.gs_add, 0, g, stream)

    # ---- device-resident throughput ------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = X.kernel_launches()
    nvl0 = nvlink_kib(local) if world > 1 else None
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    nvl1 = nvlink_kib(local) if world > 1 else None
    launches = X.kernel_launches() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public blocking gs() on host memory ----
    e2e_steps = max(1, min(args.steps, 5))
    hbuf = host.clone().pin_memory()
    X.gs(hbuf.numpy(), X.gs_double, X.gs_add, 0, g)          # warm-up (allocates staging)
    hbuf.copy_(host)
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        X.gs(hbuf.numpy(), X.gs_double, X.gs_add, 0, g)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - w0) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": n * world / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
           "ms_per_step": e2e_ms, "steps": e2e_steps,
           "api": "gs(u_host, gs_double, gs_add, 0, gsh, NULL) on pinned host memory"}

    # ---- the same call on an ordinary (pageable) host array, as a reference caller passes it
    #      (src/gs.upc:1187-1191): as it comes (the library moves it through pinned bounce rings
    #      with the host cores), and page-locked once (EXAGS_HOST_REGISTER / exags_host_register,
    #      csrc/gs_host.c) ------------------------------------------------------------------------
    pg = {}
    pbuf = np.empty(n, dtype=np.float64)
    for label, register in (("bounce_rings", False), ("registered", True)):
        if register:
            X.lib().exags_host_register.argtypes = [X.C.c_void_p, X.C.c_size_t]
            X.lib().exags_host_register.restype = X.C.c_int
            X.lib().exags_host_unregister.argtypes = [X.C.c_void_p]
            t_reg0 = time.perf_counter()
            ok_reg = X.lib().exags_host_register(pbuf.ctypes.data, pbuf.nbytes)
            t_reg = time.perf_counter() - t_reg0
        np.copyto(pbuf, host.numpy())
        X.gs(pbuf, X.gs_double, X.gs_add, 0, g)
        barrier()
        w0 = time.perf_counter()
        for _ in range(3):
            X.gs(pbuf, X.gs_double, X.gs_add, 0, g)
        barrier()
        ms = 1e3 * (time.perf_counter() - w0) / 3
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        pg[label] = {"value": n * world / (ms * 1e-3) / 1e9, "ms_per_step": ms, "steps": 3}
        if register:
            pg[label].update({"register_s": t_reg, "register_ok": bool(ok_reg)})
            X.lib().exags_host_unregister(pbuf.ctypes.data)
    del pbuf
    e2e["pageable"] = pg

    # ---- roofline of the dominant kernel (the fused interior kernel) ----
    peak, peak_src = hbm_peak()
    G, S, nsec = info["groups"], info["members"], info["sectors_f64"]
    b_alg = 2 * 32 * nsec + 4 * S + 4 * (G + 1)
    kern_ms = sum(per) / len(per)
    achieved = b_alg / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "replayed from profiles/traffic.json (one ncu --set full capture of this "
                                  "kernel on this workload), not measured in this run",
                "peak_source": peak_src,
                "algorithmic_bytes": b_alg, "kernel": "k_fused_sell<double,add,plain>",
                "kernel_ms_mean": kern_ms, "kernel_ms_min": per[0], "kernel_ms_median": per[len(per) // 2],
                "frac_vs_8TBps": achieved / 8000.0}

    check = run_check(X, g, world, n, ids_dev, ints_dev, hbuf, stream, dist)
    del ids_dev, ints_dev

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu:
            # a fresh process: the reference's UPC threads are fork()ed, which
            # must not happen inside a process that holds a CUDA context
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference",
                                    "--steps", str(args.steps), "--warmup", "2"], capture_output=True, text=True, timeout=600)
                cb = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception as exc:
                cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"cpu baseline failed: {exc}"}
        nvlink = None
        if world > 1:
            nvlink = {"counters": "NVML reports the NVLink throughput counters as not supported on this box; "
                                  "ncu's nvltx/nvlrx bytes of the send kernel are in profiles/r2_ncu_nvlink_k_pack_slots.txt",
                      "payload_bytes_sent_per_rank_per_step": 8 * info["send_slots"],
                      "peak_GBps_per_direction": 900.0}
        if nvl0 is not None and nvl1 is not None:
            # rank 0's GPU over the timed region (it also carries the barrier's few bytes of NCCL traffic);
            # the exchange occupies the link for the length of the pack kernel, a few tens of microseconds
            # per call, so the per-call average against 900 GB/s per direction is a small number by design
            tx, rx = (nvl1[0] - nvl0[0]) * 1024 / args.steps, (nvl1[1] - nvl0[1]) * 1024 / args.steps
            nvlink = {"source": "NVML field values NVLINK_THROUGHPUT_DATA_TX/RX, all links of rank 0's GPU, "
                                "difference over the timed steps",
                      "tx_bytes_per_step": tx, "rx_bytes_per_step": rx,
                      "expected_payload_bytes_per_step": 8 * info["send_slots"],
                      "avg_GBps_per_direction_over_step": tx / (ms_per_step * 1e-3) / 1e9,
                      "peak_GBps_per_direction": 900.0}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"M7 {EX}x{EY}x{EZ} hex N={ORDER} per GPU "
                                       f"({n} DOF/GPU, G={G}, S={S}), gs(add,double), "
                                       + ("1 rank" if world == 1 else
                                          f"z-slabs of {EX}x{EY}x{EZ * world}, pairwise halo exchange ({transport})"),
                           "l2": "field (805 MB/GPU) is larger than the 126 MB L2; no flush needed",
                           "setup_s": t_setup, "method": "pairwise", "exchange": transport,
                           "halo_bytes_sent_per_rank_per_step": 8 * info["send_slots"]},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "roofline": roofline, "cpu_baseline": cb, "check": check, "nvlink": nvlink}
        print(json.dumps(line), flush=True)
    X.gs_free(g)
    if dist is not None:
        dist.destroy_process_group()
    if not check["ok"]:
        raise SystemExit("bench.py: the correctness check of the timed handle failed: " + json.dumps(check))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
