"""Shared fixtures.  `-m "not gpu"` runs on a CPU box (oracle vs golden vectors, host logic, ABI surface);
`-m gpu` runs the parity tests proper on a B200, through the C ABI (ctypes -> libsmb200.so)."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count() -> int:
    try:
        rt = ctypes.CDLL("libcudart.so.12")
    except OSError:
        try:
            rt = ctypes.CDLL("libcudart.so")
        except OSError:
            return 0
    n = ctypes.c_int(0)
    if rt.cudaGetDeviceCount(ctypes.byref(n)) != 0:
        return 0
    return n.value


_N_GPUS = None


def n_gpus() -> int:
    global _N_GPUS
    if _N_GPUS is None:
        _N_GPUS = _cuda_device_count()
    return _N_GPUS


def pytest_collection_modifyitems(config, items):
    if n_gpus() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (gpu tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the oracle are built in-tree; build them if a fresh checkout lacks them."""
    lib = os.path.join(ROOT, "sparsemat_b200", "lib", "libsmb200.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    bins = [os.path.join(ROOT, "build", "replay_reference_tests"), os.path.join(ROOT, "build", "abi_example")]
    if not (os.path.exists(lib) and os.path.exists(orc) and all(os.path.exists(b) for b in bins)):
        import __graft_entry__ as ge
        ge.build()


@pytest.fixture(scope="session")
def smb():
    import sparsemat_b200
    return sparsemat_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def ctx(smb):
    c = smb.Context(0)
    yield c
    c.sync()


def rel_err(got, want, scale=None):
    """max |got - want| / scale, scale defaulting to max |want| (the tolerance base of north_star)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    s = float(np.max(np.abs(want))) if scale is None else float(scale)
    if s == 0.0:
        s = 1.0
    return float(np.max(np.abs(got - want))) / s if got.size else 0.0
