"""One rank of a multi-process test.  Launched by the tests with RANK /
WORLD_SIZE / MASTER_ADDR / MASTER_PORT / EXAGS_JOB_KEY set; ranks talk to each
other through torch.distributed (gloo on CPU, nccl not needed) for the test's
own bookkeeping, while the library under test uses its own host rendezvous
(and NCCL on GPUs).  Rank 0 checks everything against the oracle and exits
non-zero on any mismatch."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(ROOT, "exaflow-exags-upc_b200"), os.path.join(ROOT, "oracle"), HERE]

import itertools  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402


def make_ids(case, np_, rank):
    from exags_b200 import semmesh
    if case == "fixture":
        return cases.gs_test_fixture_ids(np_, rank)
    if case == "slab":
        return semmesh.box_ids(3, 3, 2, 3, ez0=2 * rank, Ez_global=2 * np_)
    if case == "bigslab":    # large enough for every chunk of the parallel setup loops to hold rows
        return semmesh.box_ids(12, 10, 3, 5, ez0=3 * rank, Ez_global=3 * np_)
    if case == "bigblocks":
        return semmesh.block_partition_ids(8, 8, 8, 4, 2, 2, 2, np_, rank)[0]
    if case == "blocks":
        return semmesh.block_partition_ids(4, 4, 4, 2, 2, 2, 2, np_, rank)[0]
    if case == "random":
        return cases.random_ids(700 + 31 * rank, 300, 99 + rank)
    if case == "sparse":     # huge, widely spaced labels: the coarse label-range filter of
        # the shared-label discovery works with a large bin width; a few labels sit
        # in bins that only one rank occupies, most are shared
        rng = np.random.default_rng(7)
        pool = np.sort(rng.choice(2**20, size=400, replace=False)).astype(np.int64) * (2**38) + 5
        mine = rng.permutation(np.concatenate([pool[rank::2], pool[::3], pool[:40]]))
        lone = (2**60 + 1000 * rank + np.arange(15)).astype(np.int64)
        return np.concatenate([mine, lone, np.zeros(3, np.int64)])
    if case == "clustered":  # each rank's labels in its own range, overlapping the next rank's by a few
        base = 10**6 * rank
        own = base + np.arange(1, 3001, dtype=np.int64)
        nxt = 10**6 * (rank + 1) + np.arange(1, 41, dtype=np.int64)      # shared with rank+1
        return np.concatenate([own, nxt, own[:500]])
    if case.startswith("fuzz"):   # fuzz<seed>: sizes, label density, zeros and flags drawn from the seed
        seed = int(case[4:])
        rng = np.random.default_rng(1000 + seed)
        nlabels = int(rng.integers(5, 400))
        n = int(rng.integers(0, 900)) if rng.random() < 0.9 else 0
        return cases.random_ids(n + 17 * rank, nlabels, 7919 * seed + rank, p_zero=float(rng.choice([0.0, 0.1, 0.5])),
                                p_flag=float(rng.choice([0.0, 0.25, 0.9])), long_group=int(rng.choice([0, 0, 30])))
    if case == "empty":      # rank 1 owns nothing
        return cases.random_ids(0 if rank == 1 else 200, 80, 5 + rank)
    raise SystemExit(f"unknown case {case}")


def gather_arrays(a, np_):
    """all ranks -> list of numpy arrays on every rank (gloo)."""
    out = [None] * np_
    dist.all_gather_object(out, a)
    return out


def main():
    case, what = sys.argv[1], sys.argv[2]
    unique = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    method = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    rank, np_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    gpu = what in ("exec", "exech", "exec2", "execw", "execg")   # exech: host pointers (staged / pipelined path)
    if not gpu:
        os.environ["EXAGS_HOST_SETUP_ONLY"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=np_)
    import exags_b200 as X
    import oracle as O
    if what == "die":
        # one rank leaves without a word; the others must end in the reference's fail()
        # ("ERROR (proc ...", exit status 1) instead of spinning in the host barrier
        X.comm_world()
        dist.barrier()
        if rank == np_ - 1:
            os._exit(7)
        X.lib().exags_comm_barrier(X.comm_world())
        X.lib().exags_comm_barrier(X.comm_world())
        sys.exit(0)
    if gpu:
        ndev = torch.cuda.device_count()
        X.lib().exags_set_device(rank % ndev)
    if rank == 0 and os.environ.get("TEST_RANK0_DELAY_S"):
        import time                       # the other ranks reach the library's rendezvous first
        time.sleep(float(os.environ["TEST_RANK0_DELAY_S"]))
    ids = make_ids(case, np_, rank)
    if unique and case.startswith("fuzz"):
        ids = np.abs(ids)     # unique on pre-flagged labels is the reference quirk of DESIGN.md 2, not compared
    g = X.gs_setup(ids, unique=unique, method=method)
    info = g.info()
    all_ids = gather_arrays(ids, np_)
    ok = True
    orc = O.Oracle(all_ids, unique=unique,
                   method=O.METHOD_ALLREDUCE if method == 3 else O.METHOD_PAIRWISE) if rank == 0 else None

    if what == "setup":
        dumps = gather_arrays([g.dump_map(w) for w in range(5)], np_)
        infos = gather_arrays(info, np_)
        if rank == 0:
            for r in range(np_):
                for w in range(5):
                    want = orc.map(r, w)
                    if not np.array_equal(dumps[r][w], want):
                        print(f"MISMATCH case={case} rank={r} map={w}\n got {dumps[r][w][:40]}\n want {want[:40]}")
                        ok = False
                if method == 2 and infos[r]["method"] != X.gs_pairwise:
                    print(f"MISMATCH crystal-router request kept method {infos[r]['method']} on rank {r}")
                    ok = False
                if infos[r]["total_shared"] != orc.total_shared:
                    print(f"MISMATCH total_shared rank={r}: {infos[r]['total_shared']} vs {orc.total_shared}")
                    ok = False
    elif what == "unique":
        mine = ids.copy()
        X.gs_unique(mine)
        got = gather_arrays(mine, np_)
        if rank == 0:
            want = [a.copy() for a in all_ids]
            O.unique(want)
            for r in range(np_):
                if not np.array_equal(got[r], want[r]):
                    print(f"MISMATCH gs_unique rank={r}\n got {got[r][:40]}\n want {want[r][:40]}")
                    ok = False
    elif what in ("exec", "exech"):
        onhost = what == "exech"
        # the maps of a handle built on the device (topology, interior/boundary
        # split) against the oracle's, like the CPU-only "setup" mode
        dumps = gather_arrays([g.dump_map(w) for w in range(5)], np_)
        if rank == 0:
            for r in range(np_):
                for w in range(5):
                    if not np.array_equal(dumps[r][w], orc.map(r, w)):
                        print(f"MISMATCH case={case} rank={r} map={w} (device-built handle)")
                        ok = False
        combos = list(itertools.product(cases.DTYPES.items(), cases.OPS, (0, 1), (0, 1, 2)))
        # gs_bpr (xxt.upc:623 calls it on sint): every method, incl. all_reduce / auto over NCCL
        combos += [(("int", np.int32), "bpr", 0, 0), (("long", np.int64), "bpr", 1, 1), (("double", np.float64), "bpr", 0, 2)]
        for (dname, dt), op, t, mode in combos:
            vn = 1 if mode == 0 else 3
            opi = cases.OPS_BPR.index(op)
            dom = X.NP_DOM[np.dtype(dt)]
            if mode == 2:
                u = [cases.values(ids.size, dt, op, 1000 * rank + k) for k in range(vn)]
                dev = [x.copy() if onhost else torch.from_numpy(x.copy()).cuda() for x in u]
                X.gs_many(dev, vn, dom, opi, t, g)
                got = dev if onhost else [d.cpu().numpy() for d in dev]
            else:
                u = cases.values(ids.size * vn, dt, op, 1000 * rank + 7)
                dev = u.copy() if onhost else torch.from_numpy(u.copy()).cuda()
                (X.gs if mode == 0 else lambda *a: X.gs_vec(a[0], vn, *a[1:]))(dev, dom, opi, t, g)
                got = dev if onhost else dev.cpu().numpy()
            ins = gather_arrays(u, np_)
            outs = gather_arrays(got, np_)
            if rank == 0:
                want = [[x.copy() for x in a] if mode == 2 else a.copy() for a in ins]
                orc.exec(want, mode, vn, O.NP_DOM[np.dtype(dt)], opi, t)
                for r in range(np_):
                    w = np.stack(want[r]) if mode == 2 else want[r]
                    gg = np.stack(outs[r]) if mode == 2 else outs[r]
                    exact = np.array_equal(gg, w)
                    if not exact and np.dtype(dt).kind == "f" and op in ("add", "mul") and method in (0, 3):
                        try:
                            np.testing.assert_array_max_ulp(gg, w, maxulp=4)
                            exact = True
                        except AssertionError:
                            pass
                    if not exact:
                        bad = np.flatnonzero(gg.ravel() != w.ravel())
                        print(f"MISMATCH exec case={case} {dname} {op} t={t} mode={mode} rank={r}: "
                              f"{bad.size} of {w.size} differ, first at {bad[:5]}")
                        ok = False
    elif what == "comm":
        # the communicator's collectives on host vectors (src/comm.upc:427-459,706-984):
        # every rank folds the ranks' vectors in rank order, so even FP sums are the
        # sequential left-to-right ones; small vectors travel through the control
        # segment, large ones through published blocks
        NPOP = {"add": np.add, "mul": np.multiply, "min": np.minimum, "max": np.maximum}
        for (dname, dt), op, vn in itertools.product(cases.DTYPES.items(), cases.OPS, (1, 5, 70001)):
            opi, dom = cases.OPS.index(op), X.NP_DOM[np.dtype(dt)]
            v = cases.values(vn, dt, op, 17 * rank + vn)
            alls = gather_arrays(v, np_)
            tot = alls[0].copy()
            for a in alls[1:]:
                tot = NPOP[op](tot, a)
            got = v.copy()
            X.comm_allreduce(got, dom, opi)
            if not np.array_equal(got, tot):
                print(f"MISMATCH comm_allreduce {dname} {op} vn={vn} rank={rank}"); ok = False
            if vn <= 5:
                excl, total = X.comm_scan(v, dom, opi)
                ident = {"add": 0, "mul": 1, "min": np.finfo(dt).max if np.dtype(dt).kind == "f" else np.iinfo(dt).max,
                         "max": -np.finfo(dt).max if np.dtype(dt).kind == "f" else np.iinfo(dt).min}[op]
                want = np.full(vn, ident, dtype=dt)
                for a in alls[:rank]:
                    want = NPOP[op](want, a)
                tot_i = np.full(vn, ident, dtype=dt)
                for a in alls:
                    tot_i = NPOP[op](tot_i, a)
                if not (np.array_equal(excl, want) and np.array_equal(total, tot_i)):
                    print(f"MISMATCH comm_scan {dname} {op} vn={vn} rank={rank}: {excl} {total} vs {want} {tot_i}"); ok = False
        rng = np.random.default_rng(3 + rank)
        a, b = rng.uniform(-1, 1, 1001), rng.uniform(-1, 1, 1001)
        local = 0.0
        for x, y in zip(a, b):
            local += x * y                                   # the reference's serial loop (src/tensor.c:12-17)
        parts = gather_arrays(local, np_)
        want = parts[0]
        for x in parts[1:]:
            want += x
        if X.comm_dot(a, b) != want:
            print(f"MISMATCH comm_dot rank={rank}"); ok = False
        for root in range(np_):
            buf = np.arange(40, dtype=np.int64) * (rank + 1) + root
            X.comm_bcast(buf, root)
            if not np.array_equal(buf, np.arange(40, dtype=np.int64) * (root + 1) + root):
                print(f"MISMATCH comm_bcast root={root} rank={rank}"); ok = False
        L = X.lib()
        import ctypes as C
        L.exags_comm_reduce__double.restype = C.c_double
        L.exags_comm_reduce__double.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint]
        got = L.exags_comm_reduce__double(C.cast(X.comm_world(), C.c_void_p), X.gs_max, a.ctypes.data, a.size)
        if got != max(gather_arrays(float(a.max()), np_)):
            print(f"MISMATCH comm_reduce_double rank={rank}"); ok = False
        okt = torch.tensor([1 if ok else 0]); dist.all_reduce(okt, op=dist.ReduceOp.MIN); ok = bool(okt.item())
    elif what == "execw":
        # weighted gs (include/exags_cuda.h): equals gs followed by u[i] *= w[i] on
        # every DOF whose label is shared with another DOF on any rank
        labs = np.abs(np.concatenate(all_ids))
        uniq, cnt = np.unique(labs[labs != 0], return_counts=True)
        shared = np.isin(np.abs(ids), uniq[cnt >= 2])
        for dt, op in itertools.product((np.float64, np.float32), cases.OPS):
            opi, dom = cases.OPS.index(op), X.NP_DOM[np.dtype(dt)]
            u = cases.values(ids.size, dt, op, 1000 * rank + 3)
            wts = np.random.default_rng(50 + rank).uniform(0.25, 1.0, ids.size).astype(dt)
            dev, dw = torch.from_numpy(u.copy()).cuda(), torch.from_numpy(wts).cuda()
            X.gs_weighted(dev, dw, dom, opi, g)
            ins, outs = gather_arrays(u, np_), gather_arrays(dev.cpu().numpy(), np_)
            ws, shs = gather_arrays(wts, np_), gather_arrays(shared, np_)
            if rank == 0:
                want = [a.copy() for a in ins]
                orc.exec(want, 0, 1, O.NP_DOM[np.dtype(dt)], opi, 0)
                for r in range(np_):
                    with np.errstate(over="ignore"):
                        w_r = np.where(shs[r], want[r] * ws[r], want[r]).astype(dt)
                    exact = np.array_equal(outs[r], w_r)
                    if not exact and op in ("add", "mul") and method in (0, 3):
                        try:
                            np.testing.assert_array_max_ulp(outs[r], w_r, maxulp=4); exact = True
                        except AssertionError:
                            pass
                    if not exact:
                        bad = np.flatnonzero(outs[r] != w_r)
                        print(f"MISMATCH execw case={case} {np.dtype(dt).name} {op} rank={r}: {bad.size} differ, first {bad[:5]}")
                        ok = False
            # a 3-component field with the same weights: equal to three plain weighted calls
            comps = [cases.values(ids.size, dt, op, 1000 * rank + 40 + k) for k in range(3)]
            want3 = []
            for c in comps:
                d3 = torch.from_numpy(c.copy()).cuda()
                X.gs_weighted(d3, dw, dom, opi, g)
                want3.append(d3)
            many = [torch.from_numpy(c.copy()).cuda() for c in comps]
            X.gs_weighted_many(many, 3, dw, dom, opi, g)
            vec = torch.from_numpy(np.stack(comps, axis=1).reshape(-1).copy()).cuda()
            X.gs_weighted_vec(vec, 3, dw, dom, opi, g)
            same = all(torch.equal(a, b) for a, b in zip(many, want3)) and \
                torch.equal(vec.view(ids.size, 3), torch.stack(want3, dim=1))
            if method in (0, 3) and op in ("add", "mul") and not same:    # all-reduce method: NCCL's association
                same = all(np.allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5) for a, b in zip(many, want3))
            if not same:
                print(f"MISMATCH execw vec/many case={case} {np.dtype(dt).name} {op} rank={rank}")
                ok = False
            okt = torch.tensor([1 if ok else 0]); dist.all_reduce(okt, op=dist.ReduceOp.MIN); ok = bool(okt.item())
    elif what == "execg":
        # np > 1 under a CUDA graph (push transport): the exchange number lives in
        # the mailbox header on the device, so the captured pack / wait / unpack
        # launches can be replayed; 2 calls captured, 3 replays, against 6 oracle calls
        for (dt, op, mode, vn) in ((np.float64, "add", 0, 1), (np.int32, "max", 1, 3)):
            dom, opi = X.NP_DOM[np.dtype(dt)], cases.OPS.index(op)
            u0 = cases.values(ids.size * vn, dt, op, 300 * rank + 11)
            dev = torch.from_numpy(u0.copy()).cuda()
            (X.gs if mode == 0 else lambda *a: X.gs_vec(a[0], vn, *a[1:]))(dev, dom, opi, 0, g)   # warm-up: mailbox, modules
            dev.copy_(torch.from_numpy(u0))
            torch.cuda.synchronize()
            dist.barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st = torch.cuda.current_stream().cuda_stream
                for _ in range(2):
                    X.gs_async(dev, mode, vn, dom, opi, 0, g, st)
            dev.copy_(torch.from_numpy(u0))
            torch.cuda.synchronize()
            dist.barrier()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            # a plain call after the replays continues the same exchange numbering
            (X.gs if mode == 0 else lambda *a: X.gs_vec(a[0], vn, *a[1:]))(dev, dom, opi, 0, g)
            ins, outs = gather_arrays(u0, np_), gather_arrays(dev.cpu().numpy(), np_)
            if rank == 0:
                want = [a.copy() for a in ins]
                for _ in range(7):
                    orc.exec(want, mode, vn, O.NP_DOM[np.dtype(dt)], opi, 0)
                for r in range(np_):
                    if not np.array_equal(outs[r], want[r]):
                        print(f"MISMATCH execg {np.dtype(dt).name} {op} mode={mode} rank={r}")
                        ok = False
            del graph
    elif what == "exec2":
        # two live handles with different neighbour sets, calls interleaved
        # (Nek keeps one handle per mesh): each has its own mailbox and epoch
        ids_b = make_ids("slab" if case != "slab" else "blocks", np_, rank)
        gb = X.gs_setup(ids_b, method=method)
        all_b = gather_arrays(ids_b, np_)
        orc_b = O.Oracle(all_b, method=O.METHOD_PAIRWISE) if rank == 0 else None
        for it in range(6):
            for (hid, hh, oo, nn) in ((0, g, orc, ids.size), (1, gb, orc_b, ids_b.size)):
                if it % 3 == 2 and hid == 0:
                    continue                      # uneven call counts per handle
                dt, op = (np.float64, "add") if (it + hid) % 2 == 0 else (np.int32, "max")
                vn = 1 + (it % 2) * 2
                u = cases.values(nn * vn, dt, op, 100 * rank + 10 * it + hid)
                dev = torch.from_numpy(u.copy()).cuda()
                dom, opi = X.NP_DOM[np.dtype(dt)], cases.OPS.index(op)
                t = it % 2
                if vn == 1:
                    X.gs(dev, dom, opi, t, hh)
                else:
                    X.gs_vec(dev, vn, dom, opi, t, hh)
                ins, outs = gather_arrays(u, np_), gather_arrays(dev.cpu().numpy(), np_)
                if rank == 0:
                    want = [a.copy() for a in ins]
                    oo.exec(want, 0 if vn == 1 else 1, vn, O.NP_DOM[np.dtype(dt)], opi, t)
                    for r in range(np_):
                        if not np.array_equal(outs[r], want[r]):
                            print(f"MISMATCH exec2 it={it} handle={hid} rank={r}")
                            ok = False
        X.gs_free(gb)
        # comm_dot / comm_allreduce on device vectors across ranks
        rng = np.random.default_rng(40 + rank)
        a, b = rng.integers(-8, 9, 50001).astype(np.float64), rng.integers(-8, 9, 50001).astype(np.float64)
        got = X.comm_dot(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        parts = gather_arrays(float(np.dot(a, b)), np_)
        if got != float(sum(parts)):
            print(f"MISMATCH comm_dot rank={rank}: {got} vs {sum(parts)}")
            ok = False
        for n_el in (5, 300000):                              # host-folded and NCCL-sized vectors
            iv = rng.integers(-1000, 1000, n_el).astype(np.int64)
            dv = torch.from_numpy(iv.copy()).cuda()
            X.comm_allreduce(dv, X.gs_long, X.gs_max)
            alls = gather_arrays(iv, np_)
            if not np.array_equal(dv.cpu().numpy(), np.max(np.stack(alls), axis=0)):
                print(f"MISMATCH device comm_allreduce n={n_el} rank={rank}")
                ok = False
        cnt = np.array([3 + rank, 10 * rank], dtype=np.int64)       # comm_scan: device input and output
        dout = torch.zeros(4, dtype=torch.int64, device="cuda")
        X.comm_scan(torch.from_numpy(cnt).cuda(), X.gs_long, X.gs_add, out=dout)
        allc = np.stack(gather_arrays(cnt, np_))
        if not np.array_equal(dout.cpu().numpy(), np.concatenate([allc[:rank].sum(axis=0), allc.sum(axis=0)])):
            print(f"MISMATCH device comm_scan rank={rank}: {dout.cpu().numpy()}")
            ok = False
        okt = torch.tensor([1 if ok else 0]); dist.all_reduce(okt, op=dist.ReduceOp.MIN); ok = bool(okt.item())
    flag = torch.tensor([0 if ok else 1])
    dist.broadcast(flag, 0)
    X.gs_free(g)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
