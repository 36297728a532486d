"""
Generates tests/golden/protocol_flow.json by RUNNING THE REFERENCE'S OWN `UsearchIndex`
(/root/reference/iscc_search/indexes/usearch/index.py, unmodified, loaded by path) through scripted
add / get / search scenarios. /root/reference does not exist on the GPU box, so the vectors are committed.

Loaded unmodified from the reference: iscc_search/schema.py, models.py, indexes/common.py, indexes/usearch/index.py,
indexes/simprint/{lmdb_ops,usearch_core,models}.py.

Stubbed, and why (none of these is installable here: no network, not in /opt/wheelhouse):
  iscc_core     -> this repo's restated codec (iscc_search_b200/iscc.py) under the names the reference calls; the codec
                   itself is pinned separately by the literal ISCCs of the reference's OpenAPI examples (tests/test_iscc_codec.py)
  lmdb          -> `FakeLmdb` below: sorted in-memory tables with py-lmdb's cursor semantics for the calls the reference
                   makes (get/put/delete/stat, cursor first/next/next_nodup/set_key/set_range/iternext/iternext_dup/
                   delete/put/putmulti, dupsort + integerdup ordering, ReadonlyError for a missing table in a read txn)
  iscc_usearch  -> `StubNphdIndex` / `StubShardedIndex128`: EXACT neighbours in (distance, key) order from
                   oracle/nphd_oracle.py, float32 distances h/n (NPHD) or h (Hamming)
  iscc_search   -> bare package object (the real __init__ needs installed distribution metadata); `iscc_search.options`
                   is the real module when it imports, else a stand-in with the same defaults
So the fixture pins the reference's HOST logic end to end (statuses, dedup, idempotent re-add, update semantics, INSTANCE
prefix rules, score aggregation, thresholds, ordering, self-exclusion, metadata enrichment, simprint scoring and grouping)
given exact neighbours; it does not pin usearch's own arithmetic (see oracle header).

    python tests/golden/make_protocol_golden.py
"""

import bisect
import importlib
import json
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
from iscc_search_b200 import iscc as codec  # noqa: E402
from oracle import nphd_oracle  # noqa: E402


# ---------------------------------------------------------------------------------------------------------------
# fake lmdb
# ---------------------------------------------------------------------------------------------------------------
class ReadonlyError(Exception):
    pass


class MapFullError(Exception):
    pass


class _Database:
    def __init__(self, name, dupsort=False, integerdup=False):
        self.name, self.dupsort, self.integerdup = name, dupsort, integerdup
        self.items = []  # sorted list of (key, sortable value, value)

    def _sv(self, value):
        return int.from_bytes(value, sys.byteorder) if self.integerdup else value

    def entry(self, key, value):
        return (key, self._sv(value), value)

    def lower(self, key):
        return bisect.bisect_left(self.items, (key,))

    def put(self, key, value, dupdata=True, overwrite=True):
        key, value = bytes(key), bytes(value)
        if self.dupsort:
            e = self.entry(key, value)
            i = bisect.bisect_left(self.items, e)
            if i < len(self.items) and self.items[i] == e:
                return False
            self.items.insert(i, e)
            return True
        i = self.lower(key)
        if i < len(self.items) and self.items[i][0] == key:
            if not overwrite:
                return False
            self.items[i] = self.entry(key, value)
            return True
        self.items.insert(i, self.entry(key, value))
        return True

    def get(self, key):
        i = self.lower(bytes(key))
        if i < len(self.items) and self.items[i][0] == bytes(key):
            return self.items[i][2]
        return None

    def delete(self, key, value=b""):
        key = bytes(key)
        if self.dupsort and value:
            e = self.entry(key, bytes(value))
            i = bisect.bisect_left(self.items, e)
            if i < len(self.items) and self.items[i] == e:
                del self.items[i]
                return True
            return False
        before = len(self.items)
        self.items = [it for it in self.items if it[0] != key]
        return len(self.items) != before


class _Cursor:
    def __init__(self, db):
        self.db, self.pos = db, None

    def _ok(self):
        return self.pos is not None and 0 <= self.pos < len(self.db.items)

    def first(self):
        self.pos = 0
        return self._ok()

    def next(self):
        self.pos = 0 if self.pos is None else self.pos + 1
        return self._ok()

    def next_nodup(self):
        if self.pos is None:
            return self.first()
        if not self._ok():
            return False
        key = self.db.items[self.pos][0]
        while self._ok() and self.db.items[self.pos][0] == key:
            self.pos += 1
        return self._ok()

    def set_key(self, key):
        i = self.db.lower(bytes(key))
        if i < len(self.db.items) and self.db.items[i][0] == bytes(key):
            self.pos = i
            return True
        self.pos = None
        return False

    def set_range(self, key):
        self.pos = self.db.lower(bytes(key))
        return self._ok()

    def key(self):
        return self.db.items[self.pos][0] if self._ok() else b""

    def value(self):
        return self.db.items[self.pos][2] if self._ok() else b""

    def item(self):
        return self.key(), self.value()

    def delete(self):
        if not self._ok():
            return False
        del self.db.items[self.pos]
        return True

    def put(self, key, value, dupdata=True, overwrite=True):
        return self.db.put(key, value, dupdata, overwrite)

    def putmulti(self, pairs, dupdata=True, overwrite=True):
        added = sum(1 for k, v in pairs if self.db.put(k, v, dupdata, overwrite))
        return len(pairs), added

    def iternext_dup(self, keys=False, values=True):
        # like py-lmdb: yields the current item and the following duplicates of its key; the cursor is left ON the last one
        if not self._ok():
            return
        key = self.db.items[self.pos][0]
        while True:
            it = self.db.items[self.pos]
            yield (it[0], it[2]) if keys and values else (it[2] if values else it[0])
            if self.pos + 1 < len(self.db.items) and self.db.items[self.pos + 1][0] == key:
                self.pos += 1
            else:
                return

    def __iter__(self):
        if self.pos is None:
            self.pos = 0
        while self._ok():
            it = self.db.items[self.pos]
            yield it[0], it[2]
            self.pos += 1

    def close(self):
        pass


class _Txn:
    def __init__(self, env, write):
        self.env, self.write = env, write

    def __enter__(self):
        if self.write:  # snapshot for rollback: an exception inside a write transaction aborts it
            self._saved = {name: list(db.items) for name, db in self.env.dbs.items()}
        return self

    def __exit__(self, exc_type, *_exc):
        if self.write and exc_type is not None:
            for name in list(self.env.dbs):
                if name in self._saved:
                    self.env.dbs[name].items = self._saved[name]
                else:
                    del self.env.dbs[name]
        return False

    def _db(self, db):
        return db if db is not None else self.env.main

    def get(self, key, default=None, db=None):
        v = self._db(db).get(key)
        return default if v is None else v

    def put(self, key, value, dupdata=True, overwrite=True, db=None):
        return self._db(db).put(key, value, dupdata, overwrite)

    def delete(self, key, value=b"", db=None):
        return self._db(db).delete(key, value)

    def stat(self, db=None):
        return {"entries": len(self._db(db).items), "psize": 4096}

    def cursor(self, db=None):
        return _Cursor(self._db(db))


class _Environment:
    def __init__(self, path, **options):
        self.path, self.options = path, options
        self.dbs, self.main = {}, _Database(None)
        self.map_size = options.get("map_size", 10485760)
        Path(path).touch()  # the manager detects an index by this file

    def open_db(self, name, txn=None, dupsort=False, dupfixed=False, integerdup=False, **_kw):
        if name not in self.dbs:
            if txn is not None and not txn.write:
                raise ReadonlyError(f"no such table in a read-only transaction: {name!r}")
            self.dbs[name] = _Database(name, dupsort, integerdup)
        return self.dbs[name]

    def begin(self, write=False, **_kw):
        return _Txn(self, write)

    def info(self):
        return {"map_size": self.map_size, "last_pgno": 1}

    def stat(self):
        return {"psize": 4096}

    def set_mapsize(self, n):
        self.map_size = n

    def close(self):
        pass


_ENVS = {}


def _lmdb_open(path, **options):
    if path not in _ENVS:  # a re-open sees the same tables (persistence across UsearchIndex instances)
        _ENVS[path] = _Environment(path, **options)
    return _ENVS[path]


# ---------------------------------------------------------------------------------------------------------------
# iscc_usearch stand-ins: exact neighbours from the oracle
# ---------------------------------------------------------------------------------------------------------------
class _Matches:
    def __init__(self, keys, distances):
        self.keys, self.distances = keys, distances

    def __len__(self):
        return len(self.keys)


_STORES = {}  # path -> StoreOracle (persistence across re-opens, like shard files)


class StubNphdIndex:
    def __init__(self, max_dim=256, path=None, **_kw):
        self.max_dim, self.path, self.dirty = max_dim, Path(path) if path else None, 0
        self.store = _STORES.setdefault(str(path), nphd_oracle.StoreOracle())
        if self.path:
            self.path.mkdir(parents=True, exist_ok=True)

    size = property(lambda self: len(self.store))
    shard_count = property(lambda self: 1 if len(self.store) else 0)
    serialized_length = property(lambda self: len(self.store) * 40)
    _active_shard_path = property(lambda self: None)

    def add(self, keys, vectors):
        self.dirty += sum(self.store.add([int(k) for k in keys], [bytes(v) for v in vectors]))

    def remove(self, keys):
        n = self.store.remove([int(k) for k in keys])
        self.dirty += n
        return n

    def __contains__(self, key):
        return int(key) in self.store

    def get(self, key):
        v = self.store.get(int(key))
        return None if v is None else np.frombuffer(v, dtype=np.uint8)

    def search(self, query, count=10, **_kw):
        hits = self.store.search(bytes(query), count, key_bytes=8)
        return _Matches(np.array([k for k, _, _ in hits], dtype=np.uint64),
                        nphd_oracle.nphd_distance_f32([h for _, h, _ in hits], [n for _, _, n in hits]) if hits else np.zeros(0, np.float32))

    def save(self):
        self.dirty = 0

    def reset(self):
        _STORES[str(self.path)] = self.store = nphd_oracle.StoreOracle()

    def drain_rotations(self):
        pass

    def close(self):
        pass


class StubShardedIndex128:
    def __init__(self, ndim=128, path=None, **_kw):
        self.ndim, self.path, self.dirty = ndim, Path(path) if path else None, 0
        self.store = _STORES.setdefault(str(path), nphd_oracle.StoreOracle())
        if self.path:
            self.path.mkdir(parents=True, exist_ok=True)

    shard_count = property(lambda self: 1 if len(self.store) else 0)
    serialized_length = property(lambda self: len(self.store) * 48)
    _active_shard_path = property(lambda self: None)

    def _normalize_batch_keys(self, keys):
        out = np.zeros(len(keys), dtype="V16")
        if len(keys):
            out.view(np.uint8).reshape(len(keys), 16)[:] = np.array([np.frombuffer(bytes(k), dtype=np.uint8) for k in keys])
        return out

    def add(self, keys, vectors):
        self.dirty += sum(self.store.add([bytes(k) for k in keys], [bytes(v) for v in vectors]))

    def remove(self, keys):
        n = self.store.remove([bytes(k) for k in keys])
        self.dirty += n
        return n

    def __contains__(self, key):
        return bytes(key) in self.store

    def __len__(self):
        return len(self.store)

    def get(self, key):
        v = self.store.get(bytes(key))
        return None if v is None else np.frombuffer(v, dtype=np.uint8)

    def search(self, vectors, count=10):
        vectors = np.atleast_2d(vectors)
        res = []
        for v in vectors:
            hits = self.store.search(bytes(v), count, key_bytes=16)
            keys = np.array([np.frombuffer(k, dtype=np.uint8) for k, _, _ in hits], dtype=np.uint8).reshape(-1, 16)
            res.append(_Matches(keys, np.array([h for _, h, _ in hits], dtype=np.float32)))
        return res[0] if len(res) == 1 else res

    def save(self):
        self.dirty = 0

    def reset(self):
        _STORES[str(self.path)] = self.store = nphd_oracle.StoreOracle()

    def drain_rotations(self):
        pass

    def close(self):
        pass


# ---------------------------------------------------------------------------------------------------------------
# module wiring
# ---------------------------------------------------------------------------------------------------------------
def load_reference():
    ic = types.ModuleType("iscc_core")
    for name in ("MT", "ST", "ST_CC", "ST_ISCC", "ST_ID", "ST_ID_REALM", "VS", "SUBTYPE_MAP", "encode_base32", "decode_base32",
                 "encode_base64", "decode_base64", "encode_header", "decode_header", "encode_length", "decode_length",
                 "encode_units", "decode_units", "iscc_clean", "gen_iscc_id"):
        setattr(ic, name, getattr(codec, name))
    ic.gen_iscc_code_v0 = codec.gen_iscc_code
    ic.gen_iscc_code = codec.gen_iscc_code
    sys.modules["iscc_core"] = ic

    lm = types.ModuleType("lmdb")
    lm.open, lm.ReadonlyError, lm.MapFullError = _lmdb_open, ReadonlyError, MapFullError
    lm._Database, lm.Environment, lm.Transaction = _Database, _Environment, _Txn
    sys.modules["lmdb"] = lm

    us = types.ModuleType("iscc_usearch")
    us.ShardedNphdIndex, us.ShardedIndex128 = StubNphdIndex, StubShardedIndex128
    sys.modules["iscc_usearch"] = us

    pkg = types.ModuleType("iscc_search")
    pkg.__path__ = [str(REF / "iscc_search")]
    pkg.dirs = types.SimpleNamespace(user_data_dir="/tmp/iscc-search-golden")
    sys.modules["iscc_search"] = pkg
    try:
        options = importlib.import_module("iscc_search.options")
    except Exception:
        options = types.ModuleType("iscc_search.options")

        class _Opts(types.SimpleNamespace):
            def override(self, update=None):
                return _Opts(**dict(self.__dict__, **(update or {})))

        options.search_opts = _Opts(match_threshold_units=0.75, match_threshold_simprints=0.75, confidence_exponent=4,
                                    oversampling_factor=20, flush_interval=100000, shard_size_units=512, shard_size_simprints=512,
                                    hnsw_expansion_add_units=128, hnsw_expansion_search_units=64, hnsw_connectivity_units=16,
                                    hnsw_expansion_add_simprints=16, hnsw_expansion_search_simprints=512,
                                    hnsw_connectivity_simprints=8)
        sys.modules["iscc_search.options"] = options
    from loguru import logger

    logger.remove()
    index_mod = importlib.import_module("iscc_search.indexes.usearch.index")
    schema = importlib.import_module("iscc_search.schema")
    return index_mod, schema


# ---------------------------------------------------------------------------------------------------------------
# scenario data
# ---------------------------------------------------------------------------------------------------------------
def unit(mt, st, body):
    return "ISCC:" + codec.encode_base32(codec.encode_header(mt, st, 0, codec.encode_length(mt, len(body) * 8)) + body)


def flip(rng, body, nbits):
    bits = np.unpackbits(np.frombuffer(body, dtype=np.uint8))
    bits[rng.choice(len(bits), size=nbits, replace=False)] ^= 1
    return np.packbits(bits).tobytes()


def rnd(rng, n):
    return bytes(rng.integers(0, 256, size=n, dtype=np.uint8))


def b64(b):
    return codec.encode_base64(b)


def make_assets(rng, n, realm=0, t0=5_000_000):
    """Families of near-duplicate assets: META / CONTENT_TEXT / DATA / INSTANCE units of mixed lengths, some simprints."""
    MT, ST_CC = codec.MT, codec.ST_CC
    n_fam = max(2, n // 4)
    fam = [{"meta": rnd(rng, 32), "content": rnd(rng, 32), "data": rnd(rng, 32), "inst": rnd(rng, 32),
            "sp64": [rnd(rng, 8) for _ in range(4)], "sp128": [rnd(rng, 16) for _ in range(3)]} for _ in range(n_fam)]
    assets = []
    for i in range(n):
        f = fam[int(rng.integers(0, n_fam))]
        ln = lambda: int(rng.choice([8, 16, 24, 32]))  # noqa: E731
        near = lambda b, m: flip(rng, b[:m], int(rng.choice([0, 0, 1, 2, 5, 9, 14, 16, 17, 20]) * m // 8)) if m else b  # noqa: E731
        units = []
        if rng.random() < 0.8:
            units.append(unit(MT.META, 0, near(f["meta"], ln())))
        if rng.random() < 0.8:
            units.append(unit(MT.CONTENT, ST_CC.TEXT, near(f["content"], ln())))
        units.append(unit(MT.DATA, 0, near(f["data"], ln())))
        inst = f["inst"] if rng.random() < 0.7 else rnd(rng, 32)
        units.append(unit(MT.INSTANCE, 0, inst[: int(rng.choice([8, 16, 32]))]))
        entry = {"iscc_id": codec.gen_iscc_id(timestamp=t0 + i * 7, hub_id=int(rng.integers(0, 4096)), realm_id=realm)["iscc"],
                 "units": units}
        if rng.random() < 0.5:
            entry["metadata"] = {"name": f"asset {i}", "source": f"https://example.com/a/{i}.txt", "n": i}
        if rng.random() < 0.6:
            sp = {}
            if rng.random() < 0.9:
                sp["CONTENT_TEXT_V0"] = [{"simprint": b64(flip(rng, s, int(rng.choice([0, 0, 1, 3, 8, 14, 16, 17, 20])))), "offset": j * 100, "size": 100}
                                         for j, s in enumerate(f["sp64"]) if rng.random() < 0.8] or [
                                             {"simprint": b64(f["sp64"][0]), "offset": 0, "size": 10}]
            if rng.random() < 0.5:
                sp["SEMANTIC_TEXT_V0"] = [{"simprint": b64(flip(rng, s, int(rng.choice([0, 2, 10, 32, 33, 40])))), "offset": j * 50, "size": 50}
                                          for j, s in enumerate(f["sp128"])]
            if sp:
                entry["simprints"] = sp
        assets.append(entry)
    return assets, fam


def dump_result(res):
    d = res.model_dump(mode="json", exclude_none=True)
    return {"query": d["query"], "global_matches": d.get("global_matches", []), "chunk_matches": d.get("chunk_matches", [])}


def run(index_mod, schema, tmp, seed=20261019, n_assets=70, **options):
    rng = np.random.default_rng(seed)
    MT, ST_CC = codec.MT, codec.ST_CC
    steps = []

    def record(op, args, fn):
        try:
            steps.append({"op": op, "args": args, "result": fn()})
        except (ValueError, FileNotFoundError, FileExistsError) as e:
            steps.append({"op": op, "args": args, "error": type(e).__name__, "message": str(e)})

    idx = index_mod.UsearchIndex(tmp / "flow", realm_id=None, max_dim=256, **options)
    assets, fam = make_assets(rng, n_assets)
    E = schema.IsccEntry

    def add(batch):
        return [r.model_dump(mode="json") for r in idx.add_assets([E(**a) for a in batch])]

    record("add_assets", {"assets": assets[:40]}, lambda: add(assets[:40]))
    # second batch: new assets, updates of existing ones (units changed, INSTANCE body changed, simprints changed),
    # the same ISCC-ID twice (last wins), and a byte-identical re-add
    upd1 = dict(assets[3], units=[unit(MT.META, 0, rnd(rng, 16)), unit(MT.DATA, 0, fam[0]["data"][:16]), unit(MT.INSTANCE, 0, rnd(rng, 16))])
    upd2 = dict(assets[5])
    upd2["simprints"] = {"CONTENT_TEXT_V0": [{"simprint": b64(fam[1]["sp64"][0]), "offset": 7, "size": 9}]}
    dup_a = dict(assets[45], metadata={"name": "first occurrence"})
    dup_b = dict(assets[45], metadata={"name": "second occurrence", "source": "https://example.com/dup"})
    batch2 = assets[40:60] + [upd1, upd2, dup_a, dup_b, assets[7]]
    record("add_assets", {"assets": batch2}, lambda: add(batch2))
    record("add_assets", {"assets": []}, lambda: add([]))
    wrong_realm = dict(assets[60], iscc_id=codec.gen_iscc_id(timestamp=9_000_000, hub_id=1, realm_id=1)["iscc"])
    record("add_assets", {"assets": [assets[60], wrong_realm]}, lambda: add([assets[60], wrong_realm]))
    record("add_assets", {"assets": [{"units": assets[61]["units"]}]}, lambda: add([{"units": assets[61]["units"]}]))
    record("add_assets", {"assets": assets[60:70]}, lambda: add(assets[60:70]))

    for a in (assets[0], assets[3], assets[45], assets[5]):
        record("get_asset", {"iscc_id": a["iscc_id"]}, lambda a=a: idx.get_asset(a["iscc_id"]).model_dump(mode="json", exclude_none=True))
    missing = codec.gen_iscc_id(timestamp=77, hub_id=7, realm_id=0)["iscc"]
    record("get_asset", {"iscc_id": missing}, lambda: idx.get_asset(missing).model_dump(mode="json"))
    record("get_asset", {"iscc_id": wrong_realm["iscc_id"]}, lambda: idx.get_asset(wrong_realm["iscc_id"]).model_dump(mode="json"))
    record("get_asset", {"iscc_id": "ISCC:NOTANID"}, lambda: idx.get_asset("ISCC:NOTANID").model_dump(mode="json"))

    Q = schema.IsccQuery

    def search(q, limit, exact=False):
        return dump_result(idx.search_assets(Q(**q), limit=limit, exact=exact))

    queries = []
    for f in fam[:10]:
        queries.append({"units": [unit(MT.META, 0, f["meta"][:8]), unit(MT.CONTENT, ST_CC.TEXT, f["content"]), unit(MT.DATA, 0, f["data"][:16]),
                                  unit(MT.INSTANCE, 0, f["inst"])]})
        queries.append({"units": [unit(MT.DATA, 0, flip(rng, f["data"], 6)), unit(MT.INSTANCE, 0, f["inst"][:8])]})
        queries.append({"units": [unit(MT.CONTENT, ST_CC.TEXT, flip(rng, f["content"][:24], 3))]})
        queries.append({"units": [unit(MT.INSTANCE, 0, f["inst"][:16])]})
    queries.append({"units": [unit(MT.CONTENT, ST_CC.IMAGE, rnd(rng, 8)), unit(MT.SEMANTIC, ST_CC.TEXT, rnd(rng, 8))]})  # unknown unit types
    queries.append({"iscc_code": codec.gen_iscc_code([unit(MT.META, 0, fam[0]["meta"][:8]), unit(MT.CONTENT, ST_CC.TEXT, fam[0]["content"][:8]),
                                                      unit(MT.DATA, 0, fam[0]["data"][:8]), unit(MT.INSTANCE, 0, fam[0]["inst"][:8])])["iscc"]})
    queries.append({"iscc_code": codec.gen_iscc_code([unit(MT.DATA, 0, fam[1]["data"][:16]), unit(MT.INSTANCE, 0, fam[1]["inst"][:16])], wide=True)["iscc"]})
    for q in queries:
        for limit in (100, 3):
            record("search_assets", {"query": q, "limit": limit}, lambda q=q, limit=limit: search(q, limit))
    for a in (assets[0], assets[3], assets[11], assets[45], assets[52]):
        record("search_assets", {"query": {"iscc_id": a["iscc_id"]}, "limit": 100}, lambda a=a: search({"iscc_id": a["iscc_id"]}, 100))
    record("search_assets", {"query": {"iscc_id": missing}, "limit": 100}, lambda: search({"iscc_id": missing}, 100))
    record("search_assets", {"query": {}, "limit": 100}, lambda: search({}, 100))

    # simprint queries: approximate (threshold search + IDF scoring) and exact (equality join + coverage x quality)
    for f in fam[:8]:
        sq = {"simprints": {"CONTENT_TEXT_V0": [b64(s) for s in f["sp64"]], "SEMANTIC_TEXT_V0": [b64(s) for s in f["sp128"][:2]]}}
        for exact in (False, True):
            for limit in (100, 2):
                record("search_assets", {"query": sq, "limit": limit, "exact": exact}, lambda sq=sq, limit=limit, exact=exact: search(sq, limit, exact))
        mixed = {"units": [unit(MT.DATA, 0, f["data"][:8])], "simprints": {"CONTENT_TEXT_V0": [b64(flip(rng, f["sp64"][1], 2))],
                                                                            "UNKNOWN_TYPE_V0": [b64(rnd(rng, 8)) + "A"]}}
        record("search_assets", {"query": mixed, "limit": 10, "exact": False}, lambda mixed=mixed: search(mixed, 10))

    # ---- second act: less common shapes
    # an asset that carries one unit type at two lengths (both share the asset's key: the last one is indexed, :420-430),
    # a query with two units of one type (per-type max, :806), an asset whose update drops a unit type and its simprints
    two_len = {"iscc_id": codec.gen_iscc_id(timestamp=8_000_001, hub_id=5, realm_id=0)["iscc"],
               "units": [unit(MT.CONTENT, ST_CC.TEXT, fam[2]["content"][:8]), unit(MT.CONTENT, ST_CC.TEXT, flip(rng, fam[2]["content"], 9)),
                         unit(MT.DATA, 0, fam[2]["data"][:24]), unit(MT.INSTANCE, 0, fam[2]["inst"][:8]), unit(MT.INSTANCE, 0, fam[3]["inst"])]}
    record("add_assets", {"assets": [two_len]}, lambda: add([two_len]))
    q_two = {"units": [unit(MT.CONTENT, ST_CC.TEXT, fam[2]["content"][:16]), unit(MT.CONTENT, ST_CC.TEXT, flip(rng, fam[2]["content"], 30)),
                       unit(MT.INSTANCE, 0, fam[3]["inst"][:16])]}
    for limit in (100, 5, 1):
        record("search_assets", {"query": q_two, "limit": limit}, lambda limit=limit: search(q_two, limit))
    with_sp = next(a for a in assets[:40] if a.get("simprints") and "CONTENT_TEXT_V0" in a["simprints"])
    stripped = {"iscc_id": with_sp["iscc_id"], "units": with_sp["units"][-2:], "metadata": {"name": "stripped"}}
    record("add_assets", {"assets": [stripped]}, lambda: add([stripped]))
    record("get_asset", {"iscc_id": with_sp["iscc_id"]}, lambda: idx.get_asset(with_sp["iscc_id"]).model_dump(mode="json", exclude_none=True))
    sp_q = {"simprints": {"CONTENT_TEXT_V0": [e["simprint"] for e in with_sp["simprints"]["CONTENT_TEXT_V0"]]}}
    for exact in (False, True):
        record("search_assets", {"query": sp_q, "limit": 50, "exact": exact}, lambda exact=exact: search(sp_q, 50, exact))
    record("search_assets", {"query": {"iscc_id": with_sp["iscc_id"]}, "limit": 50}, lambda: search({"iscc_id": with_sp["iscc_id"]}, 50))
    # the same simprints re-sent in another order (fingerprint is order independent: no-op), then with one offset changed
    sp_asset = next(a for a in assets[40:60] if a.get("simprints") and len(next(iter(a["simprints"].values()))) > 1)
    t0 = next(iter(sp_asset["simprints"]))
    reordered = dict(sp_asset, simprints={t0: list(reversed(sp_asset["simprints"][t0]))})
    record("add_assets", {"assets": [reordered]}, lambda: add([reordered]))
    moved = dict(sp_asset, simprints={t0: [dict(sp_asset["simprints"][t0][0], offset=4242)] + sp_asset["simprints"][t0][1:]})
    record("add_assets", {"assets": [moved]}, lambda: add([moved]))
    mv_q = {"simprints": {t0: [e["simprint"] for e in sp_asset["simprints"][t0]]}}
    for exact in (False, True):
        record("search_assets", {"query": mv_q, "limit": 10, "exact": exact}, lambda exact=exact: search(mv_q, 10, exact))
    # a large batch with many updates at once
    many = [dict(a, metadata={"name": f"bulk {i}", "n": i}) for i, a in enumerate(assets[10:35])]
    record("add_assets", {"assets": many}, lambda: add(many))
    for q in queries[:6]:
        record("search_assets", {"query": q, "limit": 7}, lambda q=q: search(q, 7))

    # re-open (persistence): same answers after close + open
    record("len", {}, lambda: len(idx))
    idx.close()
    idx = index_mod.UsearchIndex(tmp / "flow", max_dim=256, **options)
    record("reopen", {}, lambda: {"assets": len(idx), "realm_id": idx._realm_id})
    record("search_assets", {"query": queries[0], "limit": 100}, lambda: search(queries[0], 100))
    record("get_asset", {"iscc_id": assets[45]["iscc_id"]}, lambda: idx.get_asset(assets[45]["iscc_id"]).model_dump(mode="json", exclude_none=True))
    idx.close()
    return steps


def main():
    import shutil
    import tempfile

    index_mod, schema = load_reference()
    tmp = Path(tempfile.mkdtemp(prefix="protocol_golden_"))
    try:
        steps = run(index_mod, schema, tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    out = {"_doc": "generated by tests/golden/make_protocol_golden.py from the reference's own UsearchIndex (see its docstring)",
           "steps": steps}
    dst = Path(__file__).with_name("protocol_flow.json")
    dst.write_text(json.dumps(out, separators=(",", ":")))
    ops = {}
    for s in steps:
        ops[s["op"]] = ops.get(s["op"], 0) + 1
    errs = sum(1 for s in steps if "error" in s)
    print(f"wrote {dst} ({dst.stat().st_size} bytes): {ops}, {errs} error steps")


if __name__ == "__main__":
    main()
