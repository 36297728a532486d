"""Test configuration: `gpu` marker for tests that need a B200 (driver runs `-m gpu` on the GPU box)."""

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu` on the GPU box")


def _cuda_available():
    try:
        import ctypes

        from iscc_search_b200 import _lib

        n = ctypes.c_int()
        return _lib.lib().isx_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda():
    if not _cuda_available():
        pytest.fail("gpu test selected but no CUDA device / libisx_b200.so available (no CPU fallback exists)")
    return True


@pytest.fixture
def cpu_stores(monkeypatch):
    """Host layers above the C ABI on the oracle-backed store double (tests/fakes.py) - CPU tests of host logic only."""
    from tests import fakes

    monkeypatch.setattr("iscc_search_b200.nphd.Store", fakes.OracleStore)
    monkeypatch.setattr("iscc_search_b200.instance._default_store", lambda device: fakes.OracleStore(device, 16, 32, 0))
    return fakes.OracleStore
