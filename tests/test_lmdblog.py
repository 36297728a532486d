"""
CPU tests of the LMDB-backed host store (iscc_search_b200/lmdblog.py). The `lmdb` package is not part of this image, so it
runs on the in-memory py-lmdb model the golden generator uses for the reference itself (tests/golden/make_protocol_golden.py:
cursor, dupsort and transaction-abort semantics of the calls made), injected as the `lmdb` module.
"""

import struct
import sys
import types

import pytest

from iscc_search_b200.simprint import pack_chunk_pointer
from tests.golden import make_protocol_golden as gold


@pytest.fixture
def fake_lmdb(monkeypatch):
    gold._ENVS.clear()
    module = types.ModuleType("lmdb")
    module.open = gold._lmdb_open
    module.ReadonlyError, module.MapFullError = gold.ReadonlyError, gold.MapFullError
    monkeypatch.setitem(sys.modules, "lmdb", module)
    yield module
    gold._ENVS.clear()


def test_lmdb_log_round_trip_update_and_reference_tables(tmp_path, fake_lmdb):
    from iscc_search_b200.lmdblog import LmdbAssetLog

    log = LmdbAssetLog(tmp_path / "idx", realm_id=None, max_dim=256, lmdb_module=fake_lmdb)
    assert log.realm_id is None and len(log.assets) == 0 and list(log.simprints) == []
    log.set_realm(1)
    body = bytes(range(8))
    log.put_asset(7, b'{"a":1}')
    assert log.assets.get(7) == b'{"a":1}' and len(log.assets) == 1          # visible before the commit
    log.put_simprints("CONTENT_TEXT_V0", body, b"F" * 16, [(b"\x01" * 8, 0, 10), (b"\x02" * 8, 10, 20)])
    v0 = log.log_bytes()
    log.put_asset(7, b'{"a":2}')                                               # update
    log.put_simprints("CONTENT_TEXT_V0", body, b"G" * 16, [(b"\x03" * 8, 5, 6)])
    log.put_asset(9, b'{"b":1}')
    assert log.log_bytes() > v0
    assert dict(log.assets.items()) == {7: b'{"a":2}', 9: b'{"b":1}'}
    assert log.simprints["CONTENT_TEXT_V0"][body] == (b"G" * 16, [(b"\x03" * 8, 5, 6)])
    # the reference's own tables carry the same state: asset JSON by >Q key, chunk pointers by simprint, fingerprint by asset
    env = log.env
    with env.begin() as txn:
        assert txn.get(struct.pack(">Q", 7), db=env.open_db(b"__assets__", txn=txn)) == b'{"a":2}'
        data = env.open_db(b"__sp_CONTENT_TEXT_V0__", txn=txn, dupsort=True)
        assert [(k, v) for k, v in txn.cursor(data)] == [(b"\x03" * 8, pack_chunk_pointer(body, 5, 6))]   # old pointers are gone
        assert txn.get(body, db=env.open_db(b"__sp_assets_CONTENT_TEXT_V0__", txn=txn)) == b"G" * 16
        assert struct.unpack(">I", txn.get(b"realm_id", db=env.open_db(b"__metadata__", txn=txn)))[0] == 1
    log.close()
    again = LmdbAssetLog(tmp_path / "idx", lmdb_module=fake_lmdb)             # reopen: same content, realm from metadata
    assert again.realm_id == 1 and again.max_dim == 256 and dict(again.assets.items()) == {7: b'{"a":2}', 9: b'{"b":1}'}
    assert again.simprints.get("CONTENT_TEXT_V0").get(body)[1] == [(b"\x03" * 8, 5, 6)] and again.used_bytes() > 0
    again.close()


def test_lmdb_log_adopts_a_directory_written_by_the_reference(tmp_path, fake_lmdb):
    """Only the reference's tables exist (no __b200_sp_*): the per-asset simprint table is regrouped from them on open."""
    from iscc_search_b200.lmdblog import LmdbAssetLog

    (tmp_path / "ref").mkdir()
    env = fake_lmdb.open(str(tmp_path / "ref" / "index.lmdb"), subdir=False, max_dbs=64)
    body_a, body_b = b"A" * 8, b"B" * 8
    with env.begin(write=True) as txn:
        meta = env.open_db(b"__metadata__", txn=txn)
        txn.put(b"realm_id", struct.pack(">I", 0), db=meta)
        txn.put(b"max_dim", struct.pack(">I", 128), db=meta)
        txn.put(b"sp_types", b'["SEMANTIC_TEXT_V0"]', db=meta)
        txn.put(struct.pack(">Q", 3), b'{"x":1}', db=env.open_db(b"__assets__", txn=txn))
        data = env.open_db(b"__sp_SEMANTIC_TEXT_V0__", txn=txn, dupsort=True, dupfixed=True)
        txn.put(b"\x09" * 8, pack_chunk_pointer(body_a, 0, 4), db=data)
        txn.put(b"\x09" * 8, pack_chunk_pointer(body_b, 8, 4), db=data)
        txn.put(b"\x07" * 8, pack_chunk_pointer(body_a, 4, 4), db=data)
        txn.put(body_a, b"P" * 16, db=env.open_db(b"__sp_assets_SEMANTIC_TEXT_V0__", txn=txn))
    log = LmdbAssetLog(tmp_path / "ref", lmdb_module=fake_lmdb)
    assert log.realm_id == 0 and log.max_dim == 128 and log.assets.get(3) == b'{"x":1}'
    table = log.simprints["SEMANTIC_TEXT_V0"]
    assert sorted(table) == [body_a, body_b]
    assert table[body_a] == (b"P" * 16, [(b"\x07" * 8, 4, 4), (b"\x09" * 8, 0, 4)])   # cursor order of the dupsort table
    assert table[body_b] == (b"\x00" * 16, [(b"\x09" * 8, 8, 4)])                     # legacy marker: empty fingerprint
    log.close()


def test_protocol_flow_of_the_reference_replays_on_the_lmdb_store(tmp_path, cpu_stores, fake_lmdb):
    """The whole recorded flow of the reference's UsearchIndex (170 steps, incl. close / re-open) on the LMDB-backed log."""
    from tests.protocol_replay import replay

    counts = replay(tmp_path, asset_store="lmdb")
    assert counts["search_assets"] > 100 and counts["add_assets"] >= 6
    assert (tmp_path / "flow" / "index.lmdb").exists() and not (tmp_path / "flow" / "index.meta.json").exists()
