"""IsccIndexProtocol backend on the real HBM stores: reference-generated flow + protocol cases (needs a B200)."""

import pytest

from tests import protocol_cases
from tests.protocol_replay import replay

pytestmark = pytest.mark.gpu


def test_backend_matches_reference_flow_on_gpu(tmp_path, cuda):

    counts = replay(tmp_path)
    assert counts["search_assets"] > 100


def test_index_lifecycle_gpu(tmp_path, cuda):
    protocol_cases.case_index_lifecycle(tmp_path)


def test_add_get_search_gpu(tmp_path, cuda):
    protocol_cases.case_add_get_search(tmp_path)


def test_persistence_and_rebuild_gpu(tmp_path, cuda):
    protocol_cases.case_persistence_and_rebuild(tmp_path)


def test_concurrent_requests_share_batches_gpu(tmp_path, cuda):
    protocol_cases.case_concurrent_requests_share_batches(tmp_path)


def test_reference_index_behaviours_gpu(tmp_path, cuda):
    protocol_cases.case_reference_index_behaviours(tmp_path)


def test_multi_device_equals_single_gpu(tmp_path, cuda):
    import ctypes

    from iscc_search_b200 import _lib

    n = ctypes.c_int()
    _lib.check(_lib.lib().isx_device_count(ctypes.byref(n)))
    devices = tuple(range(min(n.value, 4))) if n.value > 1 else (0, 0)  # one GPU: two stores on the same device
    protocol_cases.case_multi_device_equals_single(tmp_path, devices=devices)


def test_crash_recovery_gpu(tmp_path, cuda):
    protocol_cases.case_crash_recovery(tmp_path)
