"""pytest configuration: markers, library loaders and shared fixtures.

`-m "not gpu"` : oracle vs compiled reference / golden vectors, host logic, C-ABI surface (no GPU).
`-m gpu`       : parity of the CUDA path against the oracle, through the C ABI (needs a B200).
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def pkg():
    try:
        return ge.load_package()
    except ImportError:
        ge.build()
        return ge.load_package()


@pytest.fixture(scope="session")
def orc_mod():
    mod = ge.load_oracle()
    if not mod.have("port"):
        ge.build()
    return mod


@pytest.fixture(scope="session")
def port(orc_mod):
    return orc_mod.Oracle("port")


@pytest.fixture(scope="session")
def ref(orc_mod):
    """The compiled reference (oracle/_ref).  Absent only if the repo was never built where
    /root/reference is mounted; tests that need it then fall back to the committed golden vectors."""
    if not orc_mod.have("ref"):
        pytest.skip("oracle/_ref/libjetpbrt_ref.so not built (needs /root/reference); golden fixtures cover this")
    return orc_mod.Oracle("ref")


@pytest.fixture(scope="session")
def checker(orc_mod):
    """Best available CPU checker: the compiled reference if present, else the pinned restatement."""
    return orc_mod.Oracle("ref") if orc_mod.have("ref") else orc_mod.Oracle("port")


@pytest.fixture(scope="session")
def gpu(pkg):
    n = pkg.device_count()
    if n <= 0:
        pytest.fail("test marked gpu but libjetpbrt_b200.so sees no CUDA device (there is no CPU fallback)")
    return n


def unit_vectors(rng, n):
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-30)
