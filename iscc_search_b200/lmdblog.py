"""
LMDB-backed host store of one index: the second implementation of the `AssetLog` interface (`assetlog.py` is the
append-log stand-in for images without the `lmdb` package - this one).

It keeps the reference's own tables in `index.lmdb` (/root/reference/iscc_search/indexes/usearch/index.py:87-103,
1165-1231), so an existing iscc-search `usearch:///` data directory opens IN PLACE and stays readable by the reference:

    __metadata__              realm_id (>I), max_dim (>I), created_at (>d), sp_types (JSON list)
    __assets__                ISCC-ID body (>Q) -> asset JSON without simprints
    __sp_{type}__   dupsort   simprint bytes -> 16-byte chunk pointer              (mirrored on every write)
    __sp_assets_{type}__      ISCC-ID body -> 16-byte fingerprint of the asset's entries

plus one table of its own per simprint type, because the backend needs an asset's entries by asset (update = drop the
old chunk pointers; the reference finds them with a scan of the dupsort table, `lmdb_ops.delete_asset_simprints`):

    __b200_sp_{type}__        ISCC-ID body -> fingerprint (16) | entries: [len u8 | simprint | offset >I | size >I] ...

It is filled from `__sp_{type}__` the first time a directory written by the reference is opened. LMDB here holds only
host-side metadata and key mapping (BASELINE north star); every vector lives in the HBM stores, which are derived data.

The `lmdb` module is injected (`lmdb_module=`) or imported; tests run it on the in-memory py-lmdb model of
tests/golden/make_protocol_golden.py, because the package is not part of this image.
"""

import json
import struct
import time
from collections.abc import Mapping
from pathlib import Path

from iscc_search_b200.simprint import pack_chunk_pointer, unpack_chunk_pointer

FINGERPRINT_BYTES = 16


def _pack_entries(fingerprint, entries):
    # type: (bytes, list[tuple[bytes, int, int]]) -> bytes
    out = [bytes(fingerprint).ljust(FINGERPRINT_BYTES, b"\0")[:FINGERPRINT_BYTES]]
    for sp_bytes, offset, size in entries:
        out.append(struct.pack(">B", len(sp_bytes)) + bytes(sp_bytes) + struct.pack(">II", offset, size))
    return b"".join(out)


def _unpack_entries(raw):
    # type: (bytes) -> tuple[bytes, list[tuple[bytes, int, int]]]
    raw = bytes(raw)
    fingerprint, pos, entries = raw[:FINGERPRINT_BYTES], FINGERPRINT_BYTES, []
    while pos < len(raw):
        n = raw[pos]
        sp_bytes = raw[pos + 1:pos + 1 + n]
        offset, size = struct.unpack(">II", raw[pos + 1 + n:pos + 9 + n])
        entries.append((sp_bytes, offset, size))
        pos += 9 + n
    return fingerprint, entries


class _AssetsView(Mapping):
    """`log.assets`: ISCC-ID key (int) -> asset JSON bytes, read through LMDB (pending writes of the open batch first)."""

    def __init__(self, log):
        self._log = log

    def __getitem__(self, key):
        value = self.get(key)
        if value is None:
            raise KeyError(key)
        return value

    def get(self, key, default=None):
        pending = self._log._pending_assets.get(key)
        if pending is not None:
            return pending
        with self._log.env.begin() as txn:
            value = txn.get(struct.pack(">Q", key), db=self._log._assets_db)
        return default if value is None else bytes(value)

    def __contains__(self, key):
        return self.get(key) is not None

    def __iter__(self):
        seen = set()
        with self._log.env.begin() as txn:
            for key_bytes, _value in txn.cursor(self._log._assets_db):
                key = struct.unpack(">Q", bytes(key_bytes))[0]
                seen.add(key)
                yield key
        for key in self._log._pending_assets:
            if key not in seen:
                yield key

    def __len__(self):
        with self._log.env.begin() as txn:
            n = txn.stat(self._log._assets_db)["entries"]
            new = sum(1 for key in self._log._pending_assets if txn.get(struct.pack(">Q", key), db=self._log._assets_db) is None)
        return n + new


class _TypeView(Mapping):
    """`log.simprints[sp_type]`: ISCC-ID body -> (fingerprint, [(simprint, offset, size), ...])."""

    def __init__(self, log, sp_type):
        self._log, self._type = log, sp_type

    def get(self, body, default=None):
        pending = self._log._pending_sp.get((self._type, bytes(body)))
        if pending is not None:
            return pending
        with self._log.env.begin() as txn:
            raw = txn.get(bytes(body), db=self._log._own_dbs[self._type])
        return default if raw is None else _unpack_entries(raw)

    def __getitem__(self, body):
        value = self.get(body)
        if value is None:
            raise KeyError(body)
        return value

    def __contains__(self, body):
        return self.get(body) is not None

    def __iter__(self):
        seen = set()
        with self._log.env.begin() as txn:
            for body, _raw in txn.cursor(self._log._own_dbs[self._type]):
                seen.add(bytes(body))
                yield bytes(body)
        for sp_type, body in self._log._pending_sp:
            if sp_type == self._type and body not in seen:
                yield body

    def __len__(self):
        return sum(1 for _ in self)


class _TablesView(Mapping):
    """`log.simprints`: simprint type -> `_TypeView`."""

    def __init__(self, log):
        self._log = log

    def __getitem__(self, sp_type):
        if sp_type not in self._log._own_dbs:
            raise KeyError(sp_type)
        return _TypeView(self._log, sp_type)

    def __iter__(self):
        return iter(list(self._log._own_dbs))

    def __len__(self):
        return len(self._log._own_dbs)


class LmdbAssetLog:
    """Same surface as `assetlog.AssetLog` (what `backend.B200Index` touches), on LMDB with the reference's tables."""

    META = "index.lmdb"   # the file whose presence marks an index directory (manager.py detects indexes the same way)

    def __init__(self, path, realm_id=None, max_dim=256, lmdb_module=None, map_size=1 << 34):
        # type: (str | Path, int | None, int, object | None, int) -> None
        if lmdb_module is None:
            import lmdb as lmdb_module  # noqa: PLC0415 - optional dependency, absent from this image
        self._lmdb = lmdb_module
        self.path = Path(path)
        self.path.mkdir(parents=True, exist_ok=True)
        self.env = lmdb_module.open(str(self.path / self.META), subdir=False, max_dbs=64, map_size=map_size)
        self._pending_assets = {}   # key -> bytes, applied by commit()
        self._pending_sp = {}       # (type, body) -> (fingerprint, entries)
        self._own_dbs, self._data_dbs, self._marks_dbs = {}, {}, {}
        self.stale_records = 0      # interface parity with AssetLog (LMDB overwrites in place)
        with self.env.begin(write=True) as txn:
            self._meta_db = self.env.open_db(b"__metadata__", txn=txn)
            self._assets_db = self.env.open_db(b"__assets__", txn=txn)
            realm_raw, dim_raw, created = (txn.get(k, db=self._meta_db) for k in (b"realm_id", b"max_dim", b"created_at"))
            self.realm_id = struct.unpack(">I", bytes(realm_raw))[0] if realm_raw is not None else realm_id
            self.max_dim = struct.unpack(">I", bytes(dim_raw))[0] if dim_raw is not None else max_dim
            self.created_at = struct.unpack(">d", bytes(created))[0] if created is not None else time.time()
            if dim_raw is None:
                txn.put(b"max_dim", struct.pack(">I", self.max_dim), db=self._meta_db)
            if created is None:
                txn.put(b"created_at", struct.pack(">d", self.created_at), db=self._meta_db)
            if realm_raw is None and self.realm_id is not None:
                txn.put(b"realm_id", struct.pack(">I", self.realm_id), db=self._meta_db)
            version = txn.get(b"b200_version", db=self._meta_db)
            self._version = struct.unpack(">Q", bytes(version))[0] if version is not None else 0
            for sp_type in self._sp_types(txn):
                self._open_type(txn, sp_type, adopt=True)
        self.assets = _AssetsView(self)
        self.simprints = _TablesView(self)

    # -- tables
    def _sp_types(self, txn):
        raw = txn.get(b"sp_types", db=self._meta_db)
        return json.loads(bytes(raw).decode()) if raw is not None else []

    def _open_type(self, txn, sp_type, adopt=False):
        if sp_type in self._own_dbs:
            return
        self._data_dbs[sp_type] = self.env.open_db(f"__sp_{sp_type}__".encode(), txn=txn, dupsort=True, dupfixed=True)
        self._marks_dbs[sp_type] = self.env.open_db(f"__sp_assets_{sp_type}__".encode(), txn=txn)
        own = self.env.open_db(f"__b200_sp_{sp_type}__".encode(), txn=txn)
        self._own_dbs[sp_type] = own
        types = self._sp_types(txn)
        if sp_type not in types:
            types.append(sp_type)
            txn.put(b"sp_types", json.dumps(types).encode(), db=self._meta_db)
        if adopt and txn.stat(own)["entries"] == 0 and txn.stat(self._data_dbs[sp_type])["entries"] > 0:
            # a directory written by the reference: regroup its simprint -> pointer table by asset, once
            per_asset = {}
            for sp_bytes, pointer in txn.cursor(self._data_dbs[sp_type]):
                body, offset, size = unpack_chunk_pointer(bytes(pointer))
                per_asset.setdefault(body, []).append((bytes(sp_bytes), offset, size))
            for body, entries in per_asset.items():
                mark = txn.get(body, db=self._marks_dbs[sp_type])
                fingerprint = bytes(mark) if mark is not None and len(mark) == FINGERPRINT_BYTES else b""
                txn.put(body, _pack_entries(fingerprint, entries), db=own)

    # -- AssetLog interface
    def set_realm(self, realm_id):
        # type: (int) -> None
        self.realm_id = realm_id
        with self.env.begin(write=True) as txn:
            txn.put(b"realm_id", struct.pack(">I", realm_id), db=self._meta_db)

    def put_asset(self, key, asset_bytes):
        # type: (int, bytes) -> None
        self._pending_assets[key] = bytes(asset_bytes)

    def put_simprints(self, sp_type, body, fingerprint, entries):
        # type: (str, bytes, bytes, list[tuple[bytes, int, int]]) -> None
        if sp_type not in self._own_dbs:
            with self.env.begin(write=True) as txn:
                self._open_type(txn, sp_type)
        self._pending_sp[(sp_type, bytes(body))] = (bytes(fingerprint), [(bytes(s), o, z) for s, o, z in entries])

    def commit(self):
        """One LMDB write transaction for everything put since the last commit (all or nothing, like the reference's batch)."""
        if not self._pending_assets and not self._pending_sp:
            return
        with self.env.begin(write=True) as txn:
            for key, asset_bytes in self._pending_assets.items():
                txn.put(struct.pack(">Q", key), asset_bytes, db=self._assets_db)
            for (sp_type, body), (fingerprint, entries) in self._pending_sp.items():
                own, data_db = self._own_dbs[sp_type], self._data_dbs[sp_type]
                old = txn.get(body, db=own)
                if old is not None:   # update: the asset's old chunk pointers leave the reference's dupsort table
                    for sp_bytes, offset, size in _unpack_entries(old)[1]:
                        txn.delete(sp_bytes, pack_chunk_pointer(body, offset, size), db=data_db)
                for sp_bytes, offset, size in entries:
                    txn.put(sp_bytes, pack_chunk_pointer(body, offset, size), dupdata=True, db=data_db)
                txn.put(body, fingerprint, db=self._marks_dbs[sp_type])
                txn.put(body, _pack_entries(fingerprint, entries), db=own)
            self._version += 1
            txn.put(b"b200_version", struct.pack(">Q", self._version), db=self._meta_db)
        self._pending_assets.clear()
        self._pending_sp.clear()

    def compact(self):
        """Nothing to do: LMDB overwrites in place."""

    def log_bytes(self):
        # type: () -> int
        """Version stamp the derived-store snapshots are tied to: the number of committed batches."""
        self.commit()
        return self._version

    def used_bytes(self):
        # type: () -> int
        return int(self.env.info()["last_pgno"] + 1) * int(self.env.stat()["psize"])

    def close(self):
        if self.env is None:
            return
        self.commit()
        self.env.close()
        self.env = None
