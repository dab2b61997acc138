import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present() -> bool:
    try:
        import ctypes
        n = ctypes.c_int(0)
        return ctypes.CDLL("libcudart.so.12").cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:  # noqa: BLE001
            return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a device: gpu-marked tests are skipped instead of failing."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product has no CPU path)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_ref():
    from oracle import oracle
    oracle.build()
    return oracle.load("ref")


@pytest.fixture(scope="session")
def oracle_fma():
    from oracle import oracle
    oracle.build()
    return oracle.load("fma")
