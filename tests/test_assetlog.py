"""Host asset log (iscc_search_b200/assetlog.py): replay, last-record-wins, torn tail, compaction, metadata."""

import json

from iscc_search_b200.assetlog import AssetLog


def _fill(log, n, tag=b"v1"):
    for i in range(n):
        log.put_asset(1000 + i, b'{"n":%d,"tag":"%s"}' % (i, tag))
    log.put_simprints("CONTENT_TEXT_V0", b"\x01" * 8, b"\xaa" * 16, [(b"\x10" * 8, 0, 10), (b"\x11" * 8, 10, 5)])
    log.commit()


def test_replay_restores_tables_and_last_record_wins(tmp_path):
    log = AssetLog(tmp_path / "idx", realm_id=None, max_dim=128)
    assert (log.realm_id, log.max_dim) == (None, 128) and log.assets == {}
    _fill(log, 5)
    log.put_asset(1002, b'{"n":2,"tag":"v2"}')
    log.put_simprints("CONTENT_TEXT_V0", b"\x01" * 8, b"\xbb" * 16, [(b"\x12" * 8, 3, 4)])
    log.set_realm(1)
    log.close()
    again = AssetLog(tmp_path / "idx", realm_id=0, max_dim=256)     # stored metadata wins over the arguments
    assert (again.realm_id, again.max_dim) == (1, 128) and len(again.assets) == 5
    assert json.loads(again.assets[1002])["tag"] == "v2" and json.loads(again.assets[1001])["tag"] == "v1"
    assert again.simprints["CONTENT_TEXT_V0"][b"\x01" * 8] == (b"\xbb" * 16, [(b"\x12" * 8, 3, 4)])
    assert again.stale_records == 2
    again.close()
    again.close()  # idempotent


def test_torn_tail_of_an_interrupted_append_is_dropped(tmp_path):
    log = AssetLog(tmp_path / "idx")
    _fill(log, 3)
    log.close()
    path = tmp_path / "idx" / AssetLog.LOG
    good = path.stat().st_size
    with open(path, "ab") as fh:
        fh.write(b"A" + (500).to_bytes(4, "big") + b"only a few bytes of the payload")
    again = AssetLog(tmp_path / "idx")
    assert len(again.assets) == 3 and path.stat().st_size == good and again.log_bytes() == good
    again.put_asset(7, b"{}")
    again.close()
    assert len(AssetLog(tmp_path / "idx").assets) == 4


def test_compaction_keeps_live_records_only(tmp_path):
    log = AssetLog(tmp_path / "idx")
    _fill(log, 4)
    for _ in range(3):
        _fill(log, 4, tag=b"again")
    before = log.log_bytes()
    assert log.stale_records == 3 * 4 + 3
    log.compact()
    assert log.stale_records == 0 and log.log_bytes() < before
    log.put_asset(99, b'{"late":true}')
    log.close()
    again = AssetLog(tmp_path / "idx")
    assert len(again.assets) == 5 and json.loads(again.assets[1003])["tag"] == "again" and again.stale_records == 0
    assert again.used_bytes() > again.log_bytes()
