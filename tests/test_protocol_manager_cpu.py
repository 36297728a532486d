"""IsccIndexProtocol cases on the oracle-backed store double: host logic of the backend without a GPU."""

from tests import protocol_cases


def test_index_lifecycle(tmp_path, cpu_stores):
    protocol_cases.case_index_lifecycle(tmp_path)


def test_add_get_search(tmp_path, cpu_stores):
    protocol_cases.case_add_get_search(tmp_path)


def test_persistence_and_rebuild(tmp_path, cpu_stores):
    protocol_cases.case_persistence_and_rebuild(tmp_path)


def test_concurrent_requests_share_batches(tmp_path, cpu_stores):
    protocol_cases.case_concurrent_requests_share_batches(tmp_path)


def test_reference_index_behaviours(tmp_path, cpu_stores):
    protocol_cases.case_reference_index_behaviours(tmp_path)


def test_multi_device_equals_single(tmp_path, cpu_stores):
    protocol_cases.case_multi_device_equals_single(tmp_path, devices=(0, 1, 2))


def test_crash_recovery(tmp_path, cpu_stores):
    protocol_cases.case_crash_recovery(tmp_path)


def test_manager_release_scratch_is_callable_on_open_indexes(cpu_stores, tmp_path):
    from iscc_search_b200 import B200IndexManager
    from iscc_search_b200.schema import IsccIndex

    m = B200IndexManager(tmp_path)
    m.create_index(IsccIndex(name="a"))
    assert m.release_scratch() == 0 and m.release_scratch("a") == 0
    import pytest

    with pytest.raises(FileNotFoundError):
        m.release_scratch("nope")
    m.close()
