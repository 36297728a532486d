"""
Protocol-level cases shared by the CPU suite (oracle-backed store double) and the GPU suite (real HBM stores).
Modelled on the reference's cross-backend conformance tests: tests/test_indexes_usearch_manager.py,
tests/test_indexes_usearch_index.py, tests/test_protocols_index.py, tests/test_server_*.py (message substrings).
"""

import numpy as np
import pytest
from pydantic import ValidationError

from iscc_search_b200 import iscc as ic
from iscc_search_b200.backend import B200IndexManager
from iscc_search_b200.schema import IsccEntry, IsccIndex, IsccQuery, Status


def unit(mt, st, body):
    return "ISCC:" + ic.encode_base32(ic.encode_header(mt, st, 0, ic.encode_length(mt, len(body) * 8)) + body)


def iscc_id(i, realm=0):
    return ic.gen_iscc_id(timestamp=1_000_000 + i, hub_id=i % 4096, realm_id=realm)["iscc"]


def flip(body, positions):
    bits = np.unpackbits(np.frombuffer(body, dtype=np.uint8))
    bits[list(positions)] ^= 1
    return np.packbits(bits).tobytes()


def rnd(seed, n):
    return bytes(np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8))


def case_index_lifecycle(tmp_path):
    from iscc_search_b200.backend import get_index

    m = get_index(f"b200://{tmp_path}?device=0", match_threshold_units=0.8)
    assert isinstance(m, B200IndexManager) and m.base_path == tmp_path and m._options == {"match_threshold_units": 0.8}
    with pytest.raises(ValueError, match="scheme"):
        get_index("usearch:///somewhere")
    assert m.list_indexes() == []
    created = m.create_index(IsccIndex(name="test"))
    assert (created.name, created.assets, created.size) == ("test", 0, 0)
    assert (tmp_path / "test").is_dir()
    with pytest.raises(FileExistsError, match="already exists"):
        m.create_index(IsccIndex(name="test"))
    with pytest.raises(ValidationError, match="String should match pattern"):
        IsccIndex(name="Invalid-Name")
    with pytest.raises(ValueError, match="Invalid index name"):
        m.create_index(IsccIndex.model_construct(name="Invalid-Name"))
    got = m.get_index("test")
    assert got.name == "test" and got.assets == 0 and got.size >= 0 and "lmdb" in got.sizes
    with pytest.raises(FileNotFoundError, match="not found"):
        m.get_index("ghost")
    m.create_index(IsccIndex(name="alpha"))
    assert [i.name for i in m.list_indexes()] == ["alpha", "test"]
    (tmp_path / "notanindex").mkdir()           # directories without the marker file and plain files are skipped
    (tmp_path / "file.txt").write_text("x")
    assert [i.name for i in m.list_indexes()] == ["alpha", "test"]
    m.delete_index("alpha")
    assert not (tmp_path / "alpha").exists()
    with pytest.raises(FileNotFoundError, match="not found"):
        m.delete_index("alpha")
    for call in (lambda: m.add_assets("ghost", []), lambda: m.get_asset("ghost", iscc_id(1)),
                 lambda: m.search_assets("ghost", IsccQuery(units=[unit(ic.MT.DATA, 0, bytes(8))])), lambda: m.rebuild("ghost")):
        with pytest.raises(FileNotFoundError, match="not found"):
            call()
    m.close()
    m.close()  # idempotent


def case_add_get_search(tmp_path):
    m = B200IndexManager(tmp_path)
    m.create_index(IsccIndex(name="test"))
    content, inst = rnd(1, 32), rnd(2, 32)
    a = IsccEntry(iscc_id=iscc_id(1), units=[unit(ic.MT.CONTENT, 0, content), unit(ic.MT.INSTANCE, 0, inst[:16])],
                  metadata={"name": "a", "source": "https://example.com/a"})
    b = IsccEntry(iscc_id=iscc_id(2), units=[unit(ic.MT.CONTENT, 0, flip(content, [3, 77, 200])), unit(ic.MT.INSTANCE, 0, rnd(3, 16))])
    c = IsccEntry(iscc_id=iscc_id(3), units=[unit(ic.MT.CONTENT, 0, rnd(4, 32)), unit(ic.MT.INSTANCE, 0, inst)])
    assert m.search_assets("test", IsccQuery(units=a.units)).global_matches == []      # empty index
    res = m.add_assets("test", [a, b, c])
    assert [(r.iscc_id, r.status) for r in res] == [(a.iscc_id, Status.created), (b.iscc_id, Status.created), (c.iscc_id, Status.created)]
    assert m.get_index("test").assets == 3
    assert m.get_asset("test", a.iscc_id) == a
    missing = iscc_id(99)
    with pytest.raises(FileNotFoundError, match="not found") as ei:
        m.get_asset("test", missing)
    assert missing in str(ei.value)
    with pytest.raises(ValueError, match="Realm mismatch"):
        m.get_asset("test", iscc_id(1, realm=1))
    with pytest.raises(ValueError, match="iscc_id"):
        m.add_assets("test", [IsccEntry(units=a.units)])
    with pytest.raises(ValueError, match="Realm ID mismatch"):
        m.add_assets("test", [IsccEntry(iscc_id=iscc_id(5, realm=1), units=a.units)])
    assert m.get_index("test").assets == 3                                            # failed batches leave nothing behind

    r = m.search_assets("test", IsccQuery(units=a.units), limit=10)
    # a and c both aggregate to 1.0 (stable sort keeps a first: it entered the aggregation first), b follows
    assert [g.iscc_id for g in r.global_matches] == [a.iscc_id, c.iscc_id, b.iscc_id]
    top = r.global_matches[0]
    assert top.score == 1.0 and top.types == {"CONTENT_TEXT_V0": 1.0, "INSTANCE_NONE_V0": 1.0} and top.metadata.name == "a"
    # b: only the CONTENT unit is near (3 of 256 bits differ); float32 NPHD turned into a Python float like the reference does
    s_b = 1.0 - float(np.float32(3) / np.float32(256))
    assert r.global_matches[2].types == {"CONTENT_TEXT_V0": s_b} and r.global_matches[2].score == s_b**4 / s_b
    # c: INSTANCE prefix relation only (128-bit stored in a, 256-bit stored in c share a prefix): binary 1.0; its random
    # CONTENT unit is reported in `types` unfiltered but does not enter the aggregate (below match_threshold_units)
    assert r.global_matches[1].types["INSTANCE_NONE_V0"] == 1.0 and r.global_matches[1].score == 1.0
    assert r.global_matches[1].types["CONTENT_TEXT_V0"] < 0.75

    # ISCC-ID query: lookup + self exclusion; unknown id -> FileNotFoundError naming the id
    r = m.search_assets("test", IsccQuery(iscc_id=a.iscc_id))
    assert [g.iscc_id for g in r.global_matches] == [c.iscc_id, b.iscc_id] and r.query.units == a.units
    with pytest.raises(FileNotFoundError) as ei:
        m.search_assets("test", IsccQuery(iscc_id=missing))
    assert missing in str(ei.value)
    with pytest.raises(ValueError, match="Query must have"):
        m.search_assets("test", IsccQuery())
    # limit
    assert len(m.search_assets("test", IsccQuery(units=a.units), limit=1).global_matches) == 1

    # update: same id, new units -> status updated, old vectors and INSTANCE rows gone
    a2 = IsccEntry(iscc_id=a.iscc_id, units=[unit(ic.MT.CONTENT, 0, rnd(7, 32)), unit(ic.MT.INSTANCE, 0, rnd(8, 16))])
    assert [r.status for r in m.add_assets("test", [a2, a2])] == [Status.updated, Status.updated]
    assert m.get_index("test").assets == 3 and m.get_asset("test", a.iscc_id) == a2
    r = m.search_assets("test", IsccQuery(units=a.units), limit=10)
    assert [g.iscc_id for g in r.global_matches][:2] == [c.iscc_id, b.iscc_id]
    assert all(g.iscc_id != a.iscc_id or g.score < 1.0 for g in r.global_matches)
    # idempotent re-add of identical bytes: still "updated", nothing changes
    assert [r.status for r in m.add_assets("test", [a2])] == [Status.updated]

    # batch front door: identical to sequential calls
    qs = [IsccQuery(units=a.units), IsccQuery(units=b.units), IsccQuery(iscc_id=c.iscc_id), IsccQuery(units=a2.units[:1])]
    assert m.search_assets_batch("test", qs, limit=10) == [m.search_assets("test", q, limit=10) for q in qs]
    m.close()


def case_persistence_and_rebuild(tmp_path):
    m = B200IndexManager(tmp_path)
    m.create_index(IsccIndex(name="keep"))
    sp = [rnd(20 + i, 8) for i in range(3)]
    entries_ = []
    for i in range(12):
        e = {"iscc_id": iscc_id(100 + i), "units": [unit(ic.MT.DATA, 0, flip(rnd(50, 16), range(i))), unit(ic.MT.INSTANCE, 0, rnd(60 + i, 8))]}
        if i % 2 == 0:
            e["simprints"] = {"CONTENT_TEXT_V0": [{"simprint": ic.encode_base64(s), "offset": 10 * j, "size": 10} for j, s in enumerate(sp)]}
        entries_.append(IsccEntry(**e))
    m.add_assets("keep", entries_)
    q_units = IsccQuery(units=[unit(ic.MT.DATA, 0, rnd(50, 16))])
    q_sp = IsccQuery(simprints={"CONTENT_TEXT_V0": [ic.encode_base64(s) for s in sp]})
    before = (m.search_assets("keep", q_units), m.search_assets("keep", q_sp))
    assert len(before[0].global_matches) == 12 and len(before[1].chunk_matches) == 6
    assert before[1].chunk_matches[0].score == 1.0 and before[1].chunk_matches[0].types["CONTENT_TEXT_V0"].matches == 3
    sizes = m.get_index("keep").sizes
    assert set(sizes) == {"lmdb", "DATA_NONE_V0", "SIMPRINT_CONTENT_TEXT_V0"}
    m.close()

    m = B200IndexManager(tmp_path)  # re-open: snapshots + log
    assert m.get_index("keep").assets == 12
    assert (m.search_assets("keep", q_units), m.search_assets("keep", q_sp)) == before
    idx = m._get_or_load_index("keep")
    assert idx.tracked_unit_types == ["DATA_NONE_V0"] and idx.tracked_simprint_types == ["CONTENT_TEXT_V0"]
    assert m.rebuild("keep") == {"unit_types": ["DATA_NONE_V0"], "simprint_types": ["CONTENT_TEXT_V0"]}
    assert m.rebuild("keep", unit_types=["CONTENT_TXT_V0"], simprint_types=["CONTENT_TXT_V0"]) == {"unit_types": [], "simprint_types": []}
    assert (m.search_assets("keep", q_units), m.search_assets("keep", q_sp)) == before
    m.close()

    # a lost snapshot is rebuilt from the asset log on open
    import shutil

    shutil.rmtree(tmp_path / "keep" / "DATA_NONE_V0")
    shutil.rmtree(tmp_path / "keep" / "SIMPRINT_CONTENT_TEXT_V0")
    m = B200IndexManager(tmp_path)
    assert (m.search_assets("keep", q_units), m.search_assets("keep", q_sp)) == before
    m.close()


def case_concurrent_requests_share_batches(tmp_path):
    """REST-style concurrency: many threads call search_assets at once; with coalesce_ms > 0 they ride shared GPU batches."""
    import threading

    m = B200IndexManager(tmp_path, coalesce_ms=20.0)
    m.create_index(IsccIndex(name="busy"))
    base = [rnd(300 + i, 16) for i in range(24)]
    m.add_assets("busy", [IsccEntry(iscc_id=iscc_id(500 + i), units=[unit(ic.MT.DATA, 0, b), unit(ic.MT.INSTANCE, 0, rnd(400 + i, 8))])
                          for i, b in enumerate(base)])
    queries = [IsccQuery(units=[unit(ic.MT.DATA, 0, flip(b, [i % 100]))]) for i, b in enumerate(base)]
    out = {}

    def worker(i):
        out[i] = m.search_assets("busy", queries[i], limit=5)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(queries))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    door = m._get_or_load_index("busy")._doors["DATA_NONE_V0"]
    assert door.requests == len(queries) and door.batches < len(queries)
    m.close()
    plain = B200IndexManager(tmp_path)   # same index, one search per request
    for i, q in enumerate(queries):
        assert out[i] == plain.search_assets("busy", q, limit=5)
        assert out[i].global_matches[0].iscc_id == iscc_id(500 + i)
    plain.close()


def case_reference_index_behaviours(tmp_path):
    """
    Behaviours the reference's own index tests pin (tests/test_indexes_usearch_index.py:141-215, 803-826, 948-981,
    984-1006, 1009-1035), restated against B200Index.
    """
    from iscc_search_b200.backend import B200Index
    from iscc_search_b200.iscc import IsccID

    idx = B200Index(tmp_path / "behaviours", realm_id=0, max_dim=256)
    # -- INSTANCE matches score 1.0 in both prefix directions (:141-215)
    base = bytes(range(1, 33))
    inst = {n: unit(ic.MT.INSTANCE, 0, base[: n // 8]) for n in (64, 128, 256)}
    content = unit(ic.MT.CONTENT, ic.ST_CC.TEXT, rnd(900, 8))
    ids = [iscc_id(700 + i) for i in range(6)]
    idx.add_assets([IsccEntry(iscc_id=ids[i], units=[inst[n], content]) for i, n in enumerate((64, 128, 256))])
    for n in (64, 128, 256):
        res = idx.search_assets(IsccQuery(units=[inst[n], content]), limit=10)
        assert {m.iscc_id for m in res.global_matches} == set(ids[:3])
        assert all(m.types["INSTANCE_NONE_V0"] == 1.0 and m.types["CONTENT_TEXT_V0"] == 1.0 and m.score == 1.0 for m in res.global_matches)
    idx.close()

    # -- matches below match_threshold_units are dropped (:803-826)
    strict = B200Index(tmp_path / "strict", realm_id=0, max_dim=256, match_threshold_units=0.99)
    body = rnd(901, 8)
    strict.add_assets([IsccEntry(iscc_id=ids[0], units=[unit(ic.MT.INSTANCE, 0, rnd(902, 16)), unit(ic.MT.CONTENT, ic.ST_CC.TEXT, body)],
                                 metadata={"name": "Test Asset"})])
    assert strict.search_assets(IsccQuery(units=[unit(ic.MT.CONTENT, ic.ST_CC.TEXT, flip(body, [5]))]), limit=10).global_matches == []  # 63/64 < 0.99
    assert len(strict.search_assets(IsccQuery(units=[unit(ic.MT.CONTENT, ic.ST_CC.TEXT, body)]), limit=10).global_matches) == 1
    strict.close()

    # -- identical re-add: "updated" but no vector mutation; a changed asset re-indexes (:948-981)
    idx = B200Index(tmp_path / "readd", realm_id=0, max_dim=256)
    c256 = unit(ic.MT.CONTENT, ic.ST_CC.TEXT, rnd(903, 32))
    i128 = unit(ic.MT.INSTANCE, 0, rnd(904, 16))
    asset = IsccEntry(iscc_id=ids[0], units=[i128, c256])
    assert idx.add_assets([asset])[0].status == Status.created
    nphd = idx._nphd_indexes["CONTENT_TEXT_V0"]
    dirty_before = nphd.dirty
    assert idx.add_assets([asset])[0].status == Status.updated and nphd.dirty == dirty_before
    changed = IsccEntry(iscc_id=ids[0], units=[i128, c256], metadata={"title": "changed"})
    assert idx.add_assets([changed])[0].status == Status.updated and nphd.dirty > dirty_before
    top = idx.search_assets(IsccQuery(units=[c256])).global_matches[0]
    assert top.iscc_id == ids[0] and top.score == 1.0

    # -- an unchanged re-add re-indexes when the derived vector went missing (:984-1006)
    nphd.remove([int(IsccID(ids[0]))])
    assert idx.search_assets(IsccQuery(units=[c256])).global_matches == []
    assert idx.add_assets([changed])[0].status == Status.updated
    top = idx.search_assets(IsccQuery(units=[c256])).global_matches[0]
    assert top.iscc_id == ids[0] and top.score == 1.0

    # -- an update drops INSTANCE bodies the asset no longer carries (:1009-1035)
    datahash = bytes(range(32))
    c = rnd(905, 32)
    idx.add_assets([IsccEntry(iscc_id=ids[1], units=[unit(ic.MT.CONTENT, ic.ST_CC.TEXT, c[:8]), unit(ic.MT.INSTANCE, 0, datahash[:8])])])
    idx.add_assets([IsccEntry(iscc_id=ids[1], units=[unit(ic.MT.CONTENT, ic.ST_CC.TEXT, c), unit(ic.MT.INSTANCE, 0, datahash)])])
    foreign = datahash[:8] + bytes(24)
    assert idx.search_assets(IsccQuery(units=[unit(ic.MT.INSTANCE, 0, foreign)])).global_matches == []
    top = idx.search_assets(IsccQuery(units=[unit(ic.MT.INSTANCE, 0, datahash)])).global_matches[0]
    assert top.iscc_id == ids[1] and top.score == 1.0
    idx.close()


def case_multi_device_equals_single(tmp_path, devices):
    """Unit stores row-sharded over several devices of one process answer exactly like a single store (order included)."""
    from iscc_search_b200.nphd import MultiDeviceNphdIndex, ShardedNphdIndex

    rng = np.random.default_rng(77)
    n = 3000
    lens = rng.choice([8, 16, 24, 32], size=n)
    base = [rnd(1000 + i, 32) for i in range(20)]
    vecs = [flip(base[i % 20], rng.choice(256, size=int(rng.integers(0, 40)), replace=False))[: lens[i]] for i in range(n)]
    keys = rng.permutation(np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    single = ShardedNphdIndex(max_dim=256)
    multi = MultiDeviceNphdIndex(max_dim=256, path=tmp_path / "multi", devices=devices)
    single.add(keys, vecs)
    multi.add(keys, vecs)
    multi.add(keys[:10], vecs[:10])  # duplicates are skipped
    assert multi.size == single.size == n and all(s.size > 0 for s in multi.shards)
    queries = [base[i % 20][: [8, 16, 24, 32][i % 4]] for i in range(24)]
    for count in (1, 10, 300):
        a, b = single.search(queries, count=count), multi.search(queries, count=count)
        assert np.array_equal(a.counts, b.counts)
        for i in range(len(queries)):
            c = int(a.counts[i])
            assert np.array_equal(a.keys[i, :c], b.keys[i, :c]) and np.array_equal(a.hamming[i, :c], b.hamming[i, :c])
            assert np.array_equal(a.nbits[i, :c], b.nbits[i, :c]) and np.array_equal(a.distances[i, :c], b.distances[i, :c])
    one_a, one_b = single.search(queries[3], count=7), multi.search(queries[3], count=7)
    assert np.array_equal(one_a.keys, one_b.keys) and np.array_equal(one_a.distances, one_b.distances)
    gone = keys[100:140]
    assert multi.remove(gone) == single.remove(gone) == 40 and multi.size == n - 40
    assert not multi.contains(gone).any() and multi.contains(keys[:5]).all() and int(keys[0]) in multi
    assert np.array_equal(multi.get(int(keys[7])), single.get(int(keys[7]))) and multi.get(int(gone[0])) is None
    multi.save()
    multi.close()
    again = MultiDeviceNphdIndex(max_dim=256, path=tmp_path / "multi", devices=devices)  # snapshots per shard
    b = again.search(queries, count=10)
    a = single.search(queries, count=10)
    assert np.array_equal(a.keys, b.keys) and np.array_equal(a.hamming, b.hamming)
    again.close()
    single.close()

    # simprint stores: threshold search with IDF scoring, equality join and document frequencies
    from iscc_search_b200.simprint import B200SimprintIndex, pack_chunk_pointer

    sp_one = B200SimprintIndex(path=None, ndim=64, oversampling_factor=20)
    sp_many = B200SimprintIndex(path=tmp_path / "sp_many", ndim=64, oversampling_factor=20, device=tuple(devices))
    sp_base = [rnd(5000 + i, 8) for i in range(12)]
    sp_keys, sp_vecs = [], []
    for a in range(60):
        body = int(2**40 + a * 7919).to_bytes(8, "big")
        for c in range(int(rng.integers(1, 6))):
            v = flip(sp_base[int(rng.integers(0, 12))], rng.choice(64, size=int(rng.choice([0, 0, 1, 3, 9, 16, 20])), replace=False))
            sp_keys.append(pack_chunk_pointer(body, c * 100, 100))
            sp_vecs.append(np.frombuffer(v, dtype=np.uint8))
    for ix in (sp_one, sp_many):
        ix.add_raw(sp_keys, sp_vecs)
    assert sp_many.size == sp_one.size == len(sp_keys) and sp_keys[5] in sp_many
    for limit, threshold in ((5, 0.0), (3, 0.75), (50, 0.9)):
        for query in (sp_base[:4], sp_base[5:6], [bytes(v) for v in sp_vecs[:3]]):
            a = sp_one.search_raw(query, limit=limit, threshold=threshold, detailed=True, doc_freq_fn="index", total_assets=60)
            b = sp_many.search_raw(query, limit=limit, threshold=threshold, detailed=True, doc_freq_fn="index", total_assets=60)
            assert a == b and (threshold > 0.8 or len(a) > 0)
            assert sp_one.search_exact(query, total_assets=60, limit=limit, threshold=threshold, detailed=True) == \
                sp_many.search_exact(query, total_assets=60, limit=limit, threshold=threshold, detailed=True)
    some = [bytes(v) for v in sp_vecs[:20]]
    assert sp_one.doc_freqs(some) == sp_many.doc_freqs(some) and sp_one.equal_keys(some, 3) == sp_many.equal_keys(some, 3)
    sp_many.remove(sp_keys[:7])
    sp_one.remove(sp_keys[:7])
    assert sp_many.size == sp_one.size and sp_keys[0] not in sp_many
    assert sp_one.search_raw(sp_base[:4], limit=5, detailed=True) == sp_many.search_raw(sp_base[:4], limit=5, detailed=True)
    sp_one.close()
    sp_many.close()

    # the protocol backend over several devices: same answers as over one
    m1 = B200IndexManager(tmp_path / "one")
    mg = B200IndexManager(tmp_path / "many", device=list(devices))
    entries_ = [IsccEntry(iscc_id=iscc_id(2000 + i), units=[unit(ic.MT.DATA, 0, vecs[i]), unit(ic.MT.INSTANCE, 0, rnd(3000 + i, 8))],
                          simprints={"CONTENT_TEXT_V0": [{"simprint": ic.encode_base64(bytes(sp_vecs[(3 * i + j) % len(sp_vecs)])), "offset": 10 * j,
                                                          "size": 10} for j in range(1 + i % 3)]} if i % 2 else None)
                for i in range(200)]
    for m in (m1, mg):
        m.create_index(IsccIndex(name="x"))
        m.add_assets("x", entries_)
    for qv in queries[:8]:
        q = IsccQuery(units=[unit(ic.MT.DATA, 0, qv)])
        assert m1.search_assets("x", q, limit=20) == mg.search_assets("x", q, limit=20)
    q = IsccQuery(simprints={"CONTENT_TEXT_V0": [ic.encode_base64(b) for b in sp_base[:5]]})
    r1, rg = m1.search_assets("x", q, limit=20), mg.search_assets("x", q, limit=20)
    assert r1 == rg and len(r1.chunk_matches) > 0
    i1, ig = m1._get_or_load_index("x"), mg._get_or_load_index("x")
    assert i1.search_assets(q, limit=20, exact=True) == ig.search_assets(q, limit=20, exact=True)
    m1.close()
    mg.close()


def case_crash_recovery(tmp_path):
    """Snapshots of the derived stores are trusted only if the asset log has not grown since they were written."""
    from iscc_search_b200.backend import B200Index

    path = tmp_path / "crashy"
    body_a, body_b = rnd(4000, 16), rnd(4001, 16)
    a = IsccEntry(iscc_id=iscc_id(4000), units=[unit(ic.MT.DATA, 0, body_a), unit(ic.MT.INSTANCE, 0, rnd(4002, 8))])
    idx = B200Index(path, realm_id=0)
    idx.add_assets([a])
    idx.close()                                     # clean shutdown: snapshot + marker
    idx = B200Index(path)
    assert idx._snapshot_sizes() == {"DATA_NONE_V0": 1}
    # an update and a new asset, then the process dies without flush/close (same row count for asset a, new vector)
    a2 = IsccEntry(iscc_id=a.iscc_id, units=[unit(ic.MT.DATA, 0, body_b), unit(ic.MT.INSTANCE, 0, rnd(4002, 8))])
    b = IsccEntry(iscc_id=iscc_id(4001), units=[unit(ic.MT.DATA, 0, flip(body_b, [1])), unit(ic.MT.INSTANCE, 0, rnd(4003, 8))])
    idx.add_assets([a2, b])
    idx._log.commit()                               # the host log reached the disk, the derived snapshots did not
    for ix in idx._nphd_indexes.values():
        ix._dirty = 0                               # "crash": nothing gets saved on the way out
    idx._nphd_indexes.clear()
    idx._log._fh.close()
    idx._log._fh = None
    idx._closed = True

    again = B200Index(path)                         # stale snapshot (old vector of a, b missing) must not be used
    res = again.search_assets(IsccQuery(units=[unit(ic.MT.DATA, 0, body_b)]), limit=10)
    assert [(m.iscc_id, m.types["DATA_NONE_V0"]) for m in res.global_matches][:2] == [(a.iscc_id, 1.0), (b.iscc_id, 1.0 - 1 / 128)]
    assert again.search_assets(IsccQuery(units=[unit(ic.MT.DATA, 0, body_a)]), limit=10).global_matches == []  # the old vector is gone
    assert again._snapshot_sizes() == {"DATA_NONE_V0": 2}
    again.close()
