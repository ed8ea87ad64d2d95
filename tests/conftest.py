"""pytest configuration: the `gpu` marker, import paths, and one-time builds.

-m "not gpu": oracle vs golden vectors, host setup logic (incl. 2-rank gloo
              runs on CPU), and the C-ABI export check.  No compute calls.
-m gpu:       parity of the CUDA path against the oracle through the C ABI.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "exaflow-exags-upc_b200")
ORACLE_DIR = os.path.join(ROOT, "oracle")
for p in (PKG_DIR, ORACLE_DIR, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")
    config.addinivalue_line("markers", "multigpu: needs at least 2 CUDA devices")


def _make(d, target=None):
    cmd = ["make", "-C", d] + ([target] if target else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"make -C {d} failed:\n{r.stdout}\n{r.stderr}")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build libexags.so and the oracle once per session (both are in-tree).
    Without a GPU (the development container) `make` always runs, so that an
    edited source can never be tested against a stale library; on a GPU box the
    prebuilt files that travelled with the snapshot are used as they are."""
    fresh = not has_gpu()
    if fresh or not os.path.exists(os.path.join(PKG_DIR, "exags_b200", "libexags.so")):
        _make(PKG_DIR)
    if fresh or not os.path.exists(os.path.join(ORACLE_DIR, "libgs_oracle.so")):
        _make(ORACLE_DIR, "libgs_oracle.so")
    return True


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def X(built):
    if not has_gpu():
        os.environ.setdefault("EXAGS_HOST_SETUP_ONLY", "1")
    import exags_b200
    return exags_b200


@pytest.fixture(scope="session")
def O(built):
    import oracle
    return oracle
