"""Multi-rank parity on real GPUs: one process per GPU, halo exchange by peer
stores over NVLink (push, the default) and by NCCL send/recv, every domain x
op x transpose x mode against the oracle (rank 0 checks).  Needs at least 2
devices; the 1-GPU box skips these and relies on the CPU multi-rank setup
tests plus the single-device exchange tests below (pull and push transports
between processes sharing GPU 0)."""
import pytest

from mp_launch import run_ranks

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ndev():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("np_,case,method", [(2, "slab", 1), (2, "fixture", 1), (2, "random", 1),
                                             (2, "slab", 3), (2, "empty", 1), (4, "blocks", 1),
                                             (4, "blocks", 3), (8, "blocks", 1)])
@pytest.mark.parametrize("transport", ["push", "nccl"])
def test_multirank_exec_nccl(built, np_, case, method, transport):
    if _ndev() < np_:
        pytest.skip(f"needs {np_} GPUs, have {_ndev()}")
    codes, outs = run_ranks(np_, [case, "exec", 0, method], timeout=600,
                            extra_env={"EXAGS_EXCHANGE": transport})
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case", [(2, "slab"), (4, "blocks"), (8, "blocks"), (3, "empty")])
def test_multirank_auto_method(built, np_, case):
    """gs_auto (what fgs_setup uses): times the transports like auto_setup,
    src/gs.upc:1113-1156, and must give the same results whichever it keeps."""
    if _ndev() < np_:
        pytest.skip(f"needs {np_} GPUs, have {_ndev()}")
    codes, outs = run_ranks(np_, [case, "exec", 0, 0], timeout=600)
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case", [(2, "slab"), (4, "blocks")])
def test_multirank_exec_unique(built, np_, case):
    if _ndev() < np_:
        pytest.skip(f"needs {np_} GPUs, have {_ndev()}")
    codes, outs = run_ranks(np_, [case, "exec", 1, 1], timeout=600)
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case,method,unique", [(2, "slab", 1, 0), (3, "random", 1, 0), (4, "blocks", 1, 0),
                                                    (2, "fixture", 3, 0), (4, "blocks", 3, 0), (2, "slab", 1, 1),
                                                    (2, "empty", 1, 0), (2, "fixture", 2, 0)])
def test_multirank_exec_one_gpu(built, np_, case, method, unique):
    """np ranks as processes sharing GPU 0: the library detects the shared
    device and exchanges halos with CUDA-IPC pulls instead of NCCL, so the
    boundary pack/unpack kernels, the interior/boundary split and the pairwise
    and all-reduce plans are all exercised on the 1-GPU box, against the oracle."""
    if _ndev() < 1:
        pytest.skip("needs a GPU")
    codes, outs = run_ranks(np_, [case, "exec", unique, method], timeout=900,
                            extra_env={"CUDA_VISIBLE_DEVICES": "0"})
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case,method,unique", [(2, "slab", 1, 0), (3, "random", 1, 0), (4, "blocks", 1, 0),
                                                    (2, "fixture", 1, 0), (2, "slab", 1, 1), (3, "empty", 1, 0),
                                                    (2, "slab", 0, 0)])
def test_multirank_exec_one_gpu_push(built, np_, case, method, unique):
    """The push transport (pack kernel stores into the neighbours' CUDA-IPC
    mapped mailboxes, flag per neighbour, one-CTA wait) between processes that
    share GPU 0: the kernels of the ranks time-slice, so this is slow but it
    runs the same code as NVLink peers do."""
    if _ndev() < 1:
        pytest.skip("needs a GPU")
    codes, outs = run_ranks(np_, [case, "exec", unique, method], timeout=900,
                            extra_env={"CUDA_VISIBLE_DEVICES": "0", "EXAGS_EXCHANGE": "push"})
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case,transport", [(2, "slab", "ipc"), (4, "blocks", "push"), (2, "random", "ipc"),
                                                 (3, "empty", "push")])
def test_multirank_host_pointers_one_gpu(built, np_, case, transport):
    """Host pointers with np > 1: symmetric maps take the chunked H2D / window
    waves / D2H pipeline with the boundary exchange after the last chunk (tiny
    chunks here so that small meshes are cut into many); flagged maps take the
    stage-run-copy path."""
    if _ndev() < 1:
        pytest.skip("needs a GPU")
    codes, outs = run_ranks(np_, [case, "exech", 0, 1], timeout=900,
                            extra_env={"CUDA_VISIBLE_DEVICES": "0", "EXAGS_EXCHANGE": transport,
                                       "EXAGS_PIPE_CHUNK_BYTES": "512"})
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case", [(2, "slab"), (4, "blocks")])
def test_multirank_host_pointers(built, np_, case):
    if _ndev() < np_:
        pytest.skip(f"needs {np_} GPUs, have {_ndev()}")
    codes, outs = run_ranks(np_, [case, "exech", 0, 1], timeout=600,
                            extra_env={"EXAGS_PIPE_CHUNK_BYTES": "512"})
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,transport", [(3, "push"), (4, "push"), (2, "ipc")])
def test_two_handles_interleaved_one_gpu(built, np_, transport):
    """Two live handles (different label sets, hence different neighbour sets
    and slot ranges) called alternately, with uneven call counts: every handle
    has its own mailbox and epoch, so their exchanges cannot mix."""
    if _ndev() < 1:
        pytest.skip("needs a GPU")
    codes, outs = run_ranks(np_, ["blocks", "exec2", 0, 1], timeout=900,
                            extra_env={"CUDA_VISIBLE_DEVICES": "0", "EXAGS_EXCHANGE": transport})
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


def test_two_handles_interleaved(built):
    if _ndev() < 4:
        pytest.skip(f"needs 4 GPUs, have {_ndev()}")
    codes, outs = run_ranks(4, ["blocks", "exec2", 0, 1], timeout=600)
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case,method,transport", [(2, "slab", 1, "ipc"), (3, "random", 1, "push"),
                                                        (4, "blocks", 1, "push"), (2, "slab", 3, "ipc")])
def test_weighted_gs_one_gpu(built, np_, case, method, transport):
    """exags_gs_weighted with np > 1 (ranks sharing GPU 0): interior groups by the
    weighted sliced / CSR kernels, boundary groups by the weighted unpack, on
    symmetric (slab, blocks) and flagged (random) maps, pairwise and all-reduce."""
    if _ndev() < 1:
        pytest.skip("needs a GPU")
    codes, outs = run_ranks(np_, [case, "execw", 0, method], timeout=900,
                            extra_env={"CUDA_VISIBLE_DEVICES": "0", "EXAGS_EXCHANGE": transport})
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case", [(2, "slab"), (4, "blocks")])
def test_weighted_gs_multi_gpu(built, np_, case):
    if _ndev() < np_:
        pytest.skip(f"needs {np_} GPUs, have {_ndev()}")
    codes, outs = run_ranks(np_, [case, "execw", 0, 1], timeout=600)
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case", [(2, "slab"), (3, "blocks")])
def test_gs_async_under_cuda_graph_one_gpu(built, np_, case):
    """exags_gs_async with np > 1 captured in a CUDA graph and replayed (push
    transport; ranks sharing GPU 0): the exchange number is device state."""
    if _ndev() < 1:
        pytest.skip("needs a GPU")
    codes, outs = run_ranks(np_, [case, "execg", 0, 1], timeout=900,
                            extra_env={"CUDA_VISIBLE_DEVICES": "0", "EXAGS_EXCHANGE": "push"})
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)


@pytest.mark.parametrize("np_,case", [(2, "slab"), (4, "blocks"), (8, "blocks")])
def test_gs_async_under_cuda_graph_multi_gpu(built, np_, case):
    if _ndev() < np_:
        pytest.skip(f"needs {np_} GPUs, have {_ndev()}")
    codes, outs = run_ranks(np_, [case, "execg", 0, 1], timeout=600)
    assert all(c == 0 for c in codes), f"exit codes {codes}\n" + "\n".join(outs)
