import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def dataset():
    """The reference's shipped `network` dataset (tests/golden/network_p3sim8.npz)."""
    z = np.load(os.path.join(GOLDEN, "network_p3sim8.npz"))
    return dict(X=z["X"], source=z["source"], target=z["target"], node_type=z["node_type"])


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "golden_ref.npz"))


@pytest.fixture(scope="session")
def legacy_xlsx():
    return np.load(os.path.join(GOLDEN, "legacy_xlsx.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def emu_lib():
    """Host (one-lane) build of the chain core, tests only."""
    import ctypes
    out_dir = os.path.join(ROOT, "tests", "emu", "_build")
    so = os.path.join(out_dir, "libbn_emu.so")
    src = os.path.join(ROOT, "tests", "emu", "host_emu.cpp")
    csrc = os.path.join(ROOT, "bayesnetworks_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w",
                               "-x", "c++", f"-I{csrc}", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.emu_score_set.restype = ctypes.c_double
    return lib


def prior_lists(source, target, n_nodes, max_par):
    """edges[tgt-1].push_back(src-1) (src/network.h:117-120) as padded arrays."""
    par = np.full((n_nodes, max_par), -1, dtype=np.int32)
    npar = np.zeros(n_nodes, dtype=np.int32)
    for s, t in zip(source, target):
        par[t - 1, npar[t - 1]] = s - 1
        npar[t - 1] += 1
    return par, npar


def centered_stats(X):
    mean = X.mean(axis=0)
    Xc = X - mean
    return mean, np.ascontiguousarray(Xc.T @ Xc)
