"""Host logic of the protocol backend against the reference-generated flow, on the oracle-backed store double (no GPU)."""

from tests.protocol_replay import replay


def test_backend_host_logic_matches_reference_flow(tmp_path, cpu_stores):
    counts = replay(tmp_path)
    assert counts["search_assets"] > 100 and counts["add_assets"] >= 6
