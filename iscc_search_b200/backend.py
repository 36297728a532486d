"""
Protocol backend: `B200Index` (one index) and `B200IndexManager` (IsccIndexProtocol).

Drop-in for the reference's usearch backend - `UsearchIndex` (/root/reference/iscc_search/indexes/usearch/index.py:87-2045)
and `UsearchIndexManager` (indexes/usearch/manager.py:26-335) - behind the same protocol
(protocols/index.py:19-174): list/create/get/delete index, add_assets, get_asset, search_assets, close, with the
same result models, status values, error types and message substrings. What changes is where the work happens:

* similarity units   -> one HBM store per unit type (`ShardedNphdIndex`, exact NPHD top-k on the GPU)
* INSTANCE units     -> `InstancePrefixIndex` (bidirectional prefix match as an `NPHD == 0` scan) instead of an LMDB dupsort table
* simprints          -> `B200SimprintIndex` per simprint type: threshold search, exact equality join and document
                        frequencies all answered by the same HBM store (the reference needs LMDB dupsort tables for the last two)
* assets / metadata  -> `AssetLog` (host; the HBM stores are derived data, rebuilt from it when a snapshot is missing or stale)

The host arithmetic that turns distances into scores and aggregates them is restated line by line from
index.py:762-881 so that results are identical given identical neighbours. `search_assets_batch` is an extension:
many queries, one GPU batch per unit type, results identical to sequential `search_assets` calls.
"""

import hashlib
import json
import shutil
import struct
import threading
from pathlib import Path

import numpy as np

from iscc_search_b200 import entries
from iscc_search_b200 import iscc as ic
from iscc_search_b200.assetlog import AssetLog
from iscc_search_b200.iscc import IsccID, IsccUnit
from iscc_search_b200.simprint import pack_chunk_pointer

DEFAULT_OPTIONS = {
    # defaults of /root/reference/iscc_search/options.py:95-125
    "match_threshold_units": 0.75,
    "match_threshold_simprints": 0.75,
    "confidence_exponent": 4,
    "oversampling_factor": 20,
    "flush_interval": 100000,
    # extension: > 0 routes the unit searches of concurrent `search_assets` calls (the REST server's thread pool,
    # server/search.py) through one `BatchingFrontDoor` per unit type, so they share GPU batches; 0 = one search per unit
    "coalesce_ms": 0.0,
    # host store of assets / metadata: "log" = assetlog.AssetLog (append log, no dependency), "lmdb" = lmdblog.LmdbAssetLog (the
    # reference's own index.lmdb tables, needs the `lmdb` package), "auto" = whatever the directory already holds, else
    # LMDB when the package is importable, else the log
    "asset_store": "auto",
}

SP_FINGERPRINT_BYTES = 16
ASSET_STORE_MARKERS = ("index.meta.json", "index.lmdb")   # AssetLog.META, LmdbAssetLog.META


def _lmdb_module(real_only=False):
    """The `lmdb` package, or None. `real_only`: ignore stand-ins that do not report a py-lmdb version (test doubles)."""
    try:
        import lmdb
    except ImportError:
        return None
    return lmdb if (not real_only or hasattr(lmdb, "__version__")) else None


def open_asset_store(path, kind="auto", realm_id=None, max_dim=256):
    # type: (str | Path, str, int | None, int) -> object
    """The host store of an index directory behind the `AssetLog` interface (see DEFAULT_OPTIONS["asset_store"])."""
    from iscc_search_b200.lmdblog import LmdbAssetLog

    path = Path(path)
    if kind not in ("auto", "log", "lmdb"):
        raise ValueError(f"asset_store must be 'auto', 'log' or 'lmdb', got {kind!r}")
    if kind == "auto":
        if (path / AssetLog.META).exists():
            kind = "log"
        elif (path / LmdbAssetLog.META).exists():
            kind = "lmdb"
        else:
            kind = "lmdb" if _lmdb_module(real_only=True) is not None else "log"
    if kind == "lmdb":
        module = _lmdb_module()
        if module is None:
            raise ImportError(f"'{path}' needs the `lmdb` package (asset_store='lmdb' / an existing index.lmdb)")
        return LmdbAssetLog(path, realm_id=realm_id, max_dim=max_dim, lmdb_module=module)
    return AssetLog(path, realm_id=realm_id, max_dim=max_dim)


def set_schema(module):
    """Use another module's result classes (inside iscc-search: `set_schema(iscc_search.schema)`)."""
    entries.schema = module


class GpuStores:
    """Factory of the HBM-resident stores of one index (the only product implementation: no CPU variant)."""

    def __init__(self, device=0):
        # `device` may be one GPU or a list of GPUs: with several, the unit stores (exact NPHD top-k) and the simprint
        # stores are row-sharded over them inside this process; the small INSTANCE store lives on the first one
        self.devices = tuple(device) if isinstance(device, (list, tuple)) else (int(device),)
        self.device = self.devices[0]

    def nphd(self, max_dim, path):
        from iscc_search_b200.nphd import MultiDeviceNphdIndex, ShardedNphdIndex

        if len(self.devices) > 1:
            return MultiDeviceNphdIndex(max_dim=max_dim, path=path, devices=self.devices)
        return ShardedNphdIndex(max_dim=max_dim, path=path, device=self.device)

    def simprint(self, path, ndim, oversampling_factor):
        from iscc_search_b200.simprint import B200SimprintIndex

        return B200SimprintIndex(path=path, ndim=ndim, oversampling_factor=oversampling_factor,
                                 device=self.devices if len(self.devices) > 1 else self.device)

    def instance(self):
        from iscc_search_b200.instance import InstancePrefixIndex

        return InstancePrefixIndex(device=self.device)


def simprint_fingerprint(simprints):
    # type: (list) -> bytes
    """Order-independent 16-byte digest of one type's (simprint, offset, size) entries - index.py:565-587."""
    triples = sorted((ic.decode_base64(sp.simprint), sp.offset, sp.size) for sp in simprints)
    hasher = hashlib.blake2b(digest_size=SP_FINGERPRINT_BYTES)
    for sp_bytes, offset, size in triples:
        hasher.update(struct.pack("!I", len(sp_bytes)))
        hasher.update(sp_bytes)
        hasher.update(struct.pack("!II", offset, size))
    return hasher.digest()


class B200Index:
    """One index directory: asset log + HBM stores. Mirrors `UsearchIndex`."""

    def __init__(self, path, realm_id=None, max_dim=256, device=0, stores=None, **options):
        # type: (str | Path, int | None, int, int, object | None, object) -> None
        unknown = set(options) - set(DEFAULT_OPTIONS)
        if unknown:
            raise ValueError(f"Unknown options: {sorted(unknown)}")
        self._opts = dict(DEFAULT_OPTIONS, **options)
        self.path = Path(path)
        self._stores = stores if stores is not None else GpuStores(device)
        self._log = open_asset_store(self.path, self._opts["asset_store"], realm_id=realm_id, max_dim=max_dim)
        self.max_dim = self._log.max_dim
        self._realm_id = self._log.realm_id
        self._nphd_indexes = {}      # unit_type -> ShardedNphdIndex
        self._simprint_indexes = {}  # sp_type -> B200SimprintIndex
        self._instance = self._stores.instance()
        self._doors = {}             # unit_type -> BatchingFrontDoor (only with coalesce_ms > 0)
        self._write_lock = threading.RLock()
        self._closed = False
        self._load_derived()

    # ---- derived stores: load snapshot, rebuild from the log when missing or stale (index.py:1602-1648, 1765-1827)
    def _unit_rows(self):
        """(unit_type -> {key: longest body}, [(key, INSTANCE body)]) from the asset log."""
        best, instance = {}, []
        for key, asset_bytes in self._log.assets.items():
            for unit_str in json.loads(asset_bytes).get("units") or []:  # plain JSON: this runs over every asset on open
                unit = IsccUnit(unit_str)
                if unit.unit_type.startswith("INSTANCE_"):
                    instance.append((key, unit.body))
                else:
                    rows = best.setdefault(unit.unit_type, {})
                    if key not in rows or len(unit.body) > len(rows[key]):
                        rows[key] = unit.body  # keep the longest body per key (index.py:1667-1685)
        return best, instance

    DERIVED_MARKER = "derived.json"

    def _snapshot_sizes(self):
        """
        {store directory name: row count} recorded at the last flush / close, or None when the asset log has grown
        since (the process died before saving): snapshots are trusted only for the log version they were written at.
        The recorded counts play the role of the reference's `nphd_count:*` / `sp_count:*` metadata (index.py:1536-1549).
        """
        try:
            marker = json.loads((self.path / self.DERIVED_MARKER).read_text())
            return marker["sizes"] if marker["log_bytes"] == self._log.log_bytes() else None
        except (OSError, ValueError, KeyError):
            return None

    def _mark_snapshots(self):
        sizes = {t: ix.size for t, ix in self._nphd_indexes.items()}
        sizes.update({f"SIMPRINT_{t}": ix.size for t, ix in self._simprint_indexes.items()})
        tmp = self.path / (self.DERIVED_MARKER + ".tmp")
        tmp.write_text(json.dumps({"log_bytes": self._log.log_bytes(), "sizes": sizes}))
        tmp.replace(self.path / self.DERIVED_MARKER)

    def _load_derived(self):
        best, instance = self._unit_rows()
        self._instance.add_many(instance)
        recorded = self._snapshot_sizes()
        rebuilt = recorded is None
        for unit_type, rows in best.items():
            index = self._stores.nphd(self.max_dim, self.path / unit_type)
            if recorded is None or index.size != recorded.get(unit_type):  # snapshot missing, stale or truncated: rebuild (index.py:1602-1648)
                index.reset()
                index.add(list(rows.keys()), list(rows.values()))
                index.save()
                rebuilt = True
            self._nphd_indexes[unit_type] = index
        for sp_type in self._log.simprints:
            keys, vectors = self._simprint_rows(sp_type)
            if not keys:
                continue
            name, ndim = f"SIMPRINT_{sp_type}", 8 * len(vectors[0])
            index = self._stores.simprint(self.path / name, ndim, self._opts["oversampling_factor"])
            if recorded is None or index.size != recorded.get(name):
                index.reset()
                index.add_raw(keys, vectors)
                index.save()
                rebuilt = True
            self._simprint_indexes[sp_type] = index
        if rebuilt and (self._nphd_indexes or self._simprint_indexes):
            self._mark_snapshots()

    @property
    def tracked_unit_types(self):
        return sorted(self._nphd_indexes)

    @property
    def tracked_simprint_types(self):
        return sorted(self._simprint_indexes)

    def rebuild(self, unit_types, simprint_types):
        # type: (list[str], list[str]) -> dict[str, list[str]]
        """Drop and rebuild the named derived stores from the asset log; types without rows are skipped (index.py:1058-1082)."""
        with self._write_lock:
            best, _instance = self._unit_rows()
            rebuilt_units, rebuilt_sp = [], []
            for unit_type in unit_types:
                rows = best.get(unit_type)
                if not rows:
                    continue
                # the old store keeps answering searches from HBM until the new one is swapped in; closing it waits for
                # the searches still running on it (isx_close takes the store's locks)
                old = self._nphd_indexes.get(unit_type)
                shutil.rmtree(self.path / unit_type, ignore_errors=True)
                index = self._stores.nphd(self.max_dim, self.path / unit_type)
                index.add(list(rows.keys()), list(rows.values()))
                index.save()
                self._nphd_indexes[unit_type] = index
                if old is not None:
                    old.close()
                rebuilt_units.append(unit_type)
            for sp_type in simprint_types:
                keys, vectors = self._simprint_rows(sp_type)
                if not keys:
                    continue
                old = self._simprint_indexes.get(sp_type)
                shutil.rmtree(self.path / f"SIMPRINT_{sp_type}", ignore_errors=True)
                index = self._stores.simprint(self.path / f"SIMPRINT_{sp_type}", 8 * len(vectors[0]), self._opts["oversampling_factor"])
                index.add_raw(keys, vectors)
                index.save()
                self._simprint_indexes[sp_type] = index
                if old is not None:
                    old.close()
                rebuilt_sp.append(sp_type)
            return {"unit_types": rebuilt_units, "simprint_types": rebuilt_sp}

    def _simprint_rows(self, sp_type):
        keys, vectors = [], []
        for body, (_f, sp_entries) in (self._log.simprints.get(sp_type) or {}).items():
            for sp_bytes, offset, size in sp_entries:
                keys.append(pack_chunk_pointer(body, offset, size))
                vectors.append(np.frombuffer(sp_bytes, dtype=np.uint8))
        return keys, vectors

    def _get_or_create_nphd_index(self, unit_type):
        if unit_type not in self._nphd_indexes:
            self._nphd_indexes[unit_type] = self._stores.nphd(self.max_dim, self.path / unit_type)
        return self._nphd_indexes[unit_type]

    def _get_or_create_simprint_index(self, sp_type, ndim):
        if sp_type not in self._simprint_indexes:
            self._simprint_indexes[sp_type] = self._stores.simprint(self.path / f"SIMPRINT_{sp_type}", ndim,
                                                                      self._opts["oversampling_factor"])
        return self._simprint_indexes[sp_type]

    # ---- add (index.py:194-537)
    def add_assets(self, assets):
        # type: (list) -> list
        if not assets:
            return []
        schema = entries.schema
        results = []
        with self._write_lock:
            # Everything that can reject an asset runs BEFORE the first mutation of the log or a store: ids, realm,
            # unit codes, simprint base64 and chunk-pointer limits are decoded here once and only the decoded values
            # are used below. The reference gets the same all-or-nothing behaviour from its LMDB write transaction.
            realm = self._realm_id
            if realm is None:  # inferred from the first asset (index.py:239-252); adopted once the batch is accepted
                if assets[0].iscc_id is None:
                    raise ValueError("Asset must have iscc_id field when adding to index")
                realm = entries.extract_realm_id(assets[0].iscc_id)
            prepared = [self._prepare_asset(asset, realm) for asset in assets]
            self._realm_id = realm
            if self._log.realm_id != self._realm_id:
                self._log.set_realm(self._realm_id)

            nphd_batches = {}     # unit_type -> ([keys], [bodies])
            nphd_updated_keys = set()
            sp_batches = {}       # sp_type -> ([chunk pointers], [vectors])
            sp_deleted_keys = {}  # sp_type -> [chunk pointers]
            instance_add, instance_remove = [], []
            last_occurrence = {asset.iscc_id: i for i, asset in enumerate(assets)}  # in-batch dedup, last wins (:263-265)
            batch_seen = set()

            for i, (asset, prep) in enumerate(zip(assets, prepared)):
                key, iscc_id_body = prep["key"], prep["body"]
                existing = self._log.assets.get(key)
                status = schema.Status.updated if (existing or key in batch_seen) else schema.Status.created
                batch_seen.add(key)
                if i != last_occurrence[asset.iscc_id]:
                    results.append(schema.IsccAddResult(iscc_id=asset.iscc_id, status=status))
                    continue

                asset_bytes = prep["bytes"]
                sp_fingerprints = prep["fingerprints"]
                # idempotent re-add: nothing to do when stored bytes, derived unit rows and simprints are all current (:294-327)
                if (existing == asset_bytes and self._nphd_units_present(key, asset.units)
                        and self._simprints_already_indexed(iscc_id_body, asset, sp_fingerprints)):
                    results.append(schema.IsccAddResult(iscc_id=asset.iscc_id, status=status))
                    continue
                if existing:
                    nphd_updated_keys.add(key)
                    new_units = set(asset.units or [])
                    for old_unit_str in entries.deserialize_asset(existing).units or []:
                        if old_unit_str in new_units:
                            continue
                        old_unit = IsccUnit(old_unit_str)
                        if old_unit.unit_type.startswith("INSTANCE_"):  # stale INSTANCE bodies go (:339-348)
                            instance_remove.append((key, old_unit.body))

                self._log.put_asset(key, asset_bytes)

                for unit_type, body in prep["units"]:
                    if unit_type.startswith("INSTANCE_"):
                        instance_add.append((key, body))
                    else:
                        batch = nphd_batches.setdefault(unit_type, ([], []))
                        batch[0].append(key)
                        batch[1].append(body)

                for sp_type, rows in prep["simprints"].items():
                    table = self._log.simprints.get(sp_type) or {}
                    if iscc_id_body in table:  # update: old chunk pointers leave the derived store (:379-383)
                        old_entries = table[iscc_id_body][1]
                        sp_deleted_keys.setdefault(sp_type, []).extend(
                            pack_chunk_pointer(iscc_id_body, o, z) for _s, o, z in old_entries)
                    batch = sp_batches.setdefault(sp_type, ([], []))
                    for sp_bytes, _offset, _size, pointer in rows:
                        batch[0].append(pointer)
                        batch[1].append(np.frombuffer(sp_bytes, dtype=np.uint8))
                    self._log.put_simprints(sp_type, iscc_id_body, sp_fingerprints[sp_type], [r[:3] for r in rows])

                results.append(schema.IsccAddResult(iscc_id=asset.iscc_id, status=status))

            self._log.commit()  # host log first, derived stores after (same order as LMDB commit -> usearch, :407-410)

            self._instance.remove_many(instance_remove)
            self._instance.add_many(instance_add)

            for unit_type, (keys, vectors) in nphd_batches.items():
                nphd_index = self._get_or_create_nphd_index(unit_type)
                if len(keys) != len(set(keys)):  # one key, two lengths of a unit type: the last one stays (:420-430)
                    unique = {}
                    for k, v in zip(keys, vectors):
                        unique[k] = v
                    keys, vectors = list(unique.keys()), list(unique.values())
                keys_to_remove = [k for k in keys if k in nphd_updated_keys]
                if keys_to_remove:
                    nphd_index.remove(keys_to_remove)
                nphd_index.add(keys, vectors)

            for sp_type, (composite_keys, sp_vectors) in sp_batches.items():
                sp_index = self._get_or_create_simprint_index(sp_type, len(sp_vectors[0]) * 8)
                if sp_type in sp_deleted_keys:
                    sp_index.remove(sp_deleted_keys[sp_type])
                sp_index.add_raw(composite_keys, sp_vectors)

            flush_interval = self._opts["flush_interval"]
            if flush_interval > 0:
                for index in list(self._nphd_indexes.values()) + list(self._simprint_indexes.values()):
                    if index.dirty >= flush_interval:
                        index.save()
        return results

    def _prepare_asset(self, asset, realm):
        # type: (object, int) -> dict
        """Decode one incoming asset completely (raises ValueError on anything malformed); nothing is stored here."""
        if asset.iscc_id is None:
            raise ValueError("Asset must have iscc_id field when adding to index")
        asset_realm = entries.extract_realm_id(asset.iscc_id)
        if realm != asset_realm:
            raise ValueError(
                f"Realm ID mismatch: index has realm={realm}, "
                f"but asset '{asset.iscc_id}' has realm={asset_realm}. "
                f"All assets in an index must have the same realm ID."
            )
        iscc_id_obj = IsccID(asset.iscc_id)
        body = iscc_id_obj.body
        units = []
        for unit_str in asset.units or []:
            unit = IsccUnit(unit_str)
            units.append((unit.unit_type, unit.body))
        simprints, fingerprints = {}, {}
        for sp_type, sp_list in (asset.simprints or {}).items():
            rows = []
            for sp_obj in sp_list:
                sp_bytes = ic.decode_base64(sp_obj.simprint)
                rows.append((sp_bytes, sp_obj.offset, sp_obj.size, pack_chunk_pointer(body, sp_obj.offset, sp_obj.size)))
            if rows and len({len(r[0]) for r in rows}) != 1:
                raise ValueError(f"Simprints of type '{sp_type}' in asset '{asset.iscc_id}' have different lengths")
            simprints[sp_type] = rows
            fingerprints[sp_type] = simprint_fingerprint(sp_list)
        return {"key": int(iscc_id_obj), "body": body, "bytes": entries.serialize_asset(asset), "units": units,
                "simprints": simprints, "fingerprints": fingerprints}

    def _nphd_units_present(self, key, units):
        for unit_str in units or []:
            unit_type = IsccUnit(unit_str).unit_type
            if unit_type.startswith("INSTANCE_"):
                continue
            nphd_index = self._nphd_indexes.get(unit_type)
            if nphd_index is None or key not in nphd_index:
                return False
        return True

    def _simprints_already_indexed(self, iscc_id_body, asset, fingerprints):
        """Subset semantics of index.py:589-655: every incoming type is stored with the same fingerprint and its rows exist."""
        for sp_type, sp_list in (asset.simprints or {}).items():
            stored = (self._log.simprints.get(sp_type) or {}).get(iscc_id_body)
            if stored is None or stored[0] != fingerprints[sp_type]:
                return False
            sp_index = self._simprint_indexes.get(sp_type)
            if sp_index is None:
                return False
            for sp_obj in sp_list:
                if pack_chunk_pointer(iscc_id_body, sp_obj.offset, sp_obj.size) not in sp_index:
                    return False
        return True

    # ---- get (index.py:704-733)
    def get_asset(self, iscc_id):
        entries.validate_iscc_id(iscc_id, expected_realm=self._realm_id)
        asset_bytes = self._log.assets.get(int(IsccID(iscc_id)))
        if asset_bytes is None:
            raise FileNotFoundError(f"Asset '{iscc_id}' not found in index")
        return entries.deserialize_asset(asset_bytes)

    # ---- search (index.py:735-881)
    def search_assets(self, query, limit=100, exact=False):
        return self.search_assets_batch([query], limit, exact)[0]

    def release_scratch(self):
        # type: () -> int
        """Free the per-search working memory of every derived store of this index (rows stay); returns the bytes freed."""
        stores = list(self._nphd_indexes.values()) + list(self._simprint_indexes.values()) + [self._instance]
        return sum(store.release_scratch() for store in stores)

    def search_assets_batch(self, queries, limit=100, exact=False):
        # type: (list, int, bool) -> list
        """Many queries at once: one GPU batch per unit type; each result equals `search_assets(query)`."""
        try:
            return self._search_assets_batch(queries, limit, exact)
        finally:
            if len(queries) >= 1024:   # a large batch grows GB of candidate lists per store: give them back
                self.release_scratch()

    def _search_assets_batch(self, queries, limit=100, exact=False):
        schema = entries.schema
        prepared = []
        for query in queries:
            query_iscc_id = None
            if query.iscc_id:  # look-up + self-exclusion (:762-769)
                query_iscc_id = query.iscc_id
                asset = self.get_asset(query.iscc_id)
                query = schema.IsccQuery(iscc_code=asset.iscc_code, units=asset.units, simprints=_as_query_simprints(asset.simprints))
            prepared.append((entries.normalize_query(query), query_iscc_id))

        # similarity units of all queries, grouped per unit type -> one batched exact NPHD top-k each
        wanted = {}  # unit_type -> [(query index, body)]
        for qi, (query, _qid) in enumerate(prepared):
            for unit_str in query.units or []:
                unit = IsccUnit(unit_str)
                if not unit.unit_type.startswith("INSTANCE_") and unit.unit_type in self._nphd_indexes:
                    wanted.setdefault(unit.unit_type, []).append((qi, unit.body))
        similar = {}  # (query index, unit_type, body) -> {key: score}
        for unit_type, items in wanted.items():
            bodies = list(dict.fromkeys(body for _qi, body in items))
            for body, scores in zip(bodies, self._search_similarity_units(unit_type, bodies, limit)):
                for qi, b in items:
                    if b == body:
                        similar[(qi, unit_type, body)] = scores

        out = []
        for qi, (query, query_iscc_id) in enumerate(prepared):
            chunk_matches = []
            if self._simprint_indexes and query.simprints:
                chunk_matches = self._search_simprints(query, limit, exact=exact)
            matches = []
            if query.units:
                aggregated = {}  # key -> {unit_type: score}
                for unit_str in query.units:
                    unit = IsccUnit(unit_str)
                    unit_type = unit.unit_type
                    if unit_type.startswith("INSTANCE_"):
                        for key, score in self._instance.search(unit.body).items():
                            aggregated.setdefault(key, {})[unit_type] = score
                    elif unit_type in self._nphd_indexes:
                        for key, score in similar[(qi, unit_type, unit.body)].items():
                            per_type = aggregated.setdefault(key, {})
                            per_type[unit_type] = max(per_type.get(unit_type, 0.0), score)

                threshold, exponent = self._opts["match_threshold_units"], self._opts["confidence_exponent"]
                scored_results = []
                for key, unit_scores in aggregated.items():
                    confident = {t: s for t, s in unit_scores.items() if s >= threshold}
                    if not confident:
                        continue
                    weighted_sum = sum(s**exponent for s in confident.values())
                    weight_sum = sum(s for s in confident.values())
                    total_score = weighted_sum / weight_sum if weight_sum > 0 else 0.0
                    scored_results.append((key, total_score, unit_scores))
                if query_iscc_id:
                    query_key = int(IsccID(query_iscc_id))
                    scored_results = [r for r in scored_results if r[0] != query_key]
                scored_results.sort(key=lambda x: x[1], reverse=True)
                scored_results = scored_results[:limit]
                for key, total_score, unit_scores in scored_results:
                    metadata = None
                    asset_bytes = self._log.assets.get(key)
                    if asset_bytes is not None:
                        asset = entries.deserialize_asset(asset_bytes)
                        if asset.metadata:
                            metadata = asset.metadata
                    matches.append(schema.IsccGlobalMatch(iscc_id=str(IsccID.from_int(key, self._realm_id)), score=total_score,
                                                          types=unit_scores, metadata=metadata))
            if query_iscc_id:
                chunk_matches = [m for m in chunk_matches if m.iscc_id != query_iscc_id]
            out.append(schema.IsccSearchResult(query=query, global_matches=matches, chunk_matches=chunk_matches))
        return out

    def _search_similarity_units(self, unit_type, bodies, limit):
        # type: (str, list[bytes], int) -> list[dict[int, float]]
        """Batched `_search_similarity_unit` (index.py:2024-2045): score = max(0, 1 - float(float32 NPHD))."""
        nphd_index = self._nphd_indexes[unit_type]
        if nphd_index.size == 0:
            return [{} for _ in bodies]
        count = min(int(limit), getattr(nphd_index, "max_count", int(limit)))  # REST `limit` is unbounded above (server/search.py:22)
        if len(bodies) == 1 and self._opts["coalesce_ms"] > 0:
            per_query = [self._door(unit_type, nphd_index).search(np.frombuffer(bodies[0], dtype=np.uint8), count=count)]
        else:
            res = nphd_index.search([np.frombuffer(b, dtype=np.uint8) for b in bodies], count=count)
            per_query = [res] if len(bodies) == 1 else [res[i] for i in range(len(bodies))]
        out = []
        for m in per_query:
            scores = {}
            for key, distance in zip(m.keys, m.distances):
                scores[int(key)] = max(0.0, 1.0 - float(distance))
            out.append(scores)
        return out

    def _door(self, unit_type, nphd_index):
        door = self._doors.get(unit_type)
        if door is None or door.index is not nphd_index:
            with self._write_lock:
                door = self._doors.get(unit_type)
                if door is None or door.index is not nphd_index:
                    from iscc_search_b200.frontdoor import BatchingFrontDoor

                    if door is not None:
                        door.close()
                    door = self._doors[unit_type] = BatchingFrontDoor(nphd_index, max_delay_ms=self._opts["coalesce_ms"])
        return door

    def _search_instance_unit(self, instance_code):
        return self._instance.search(instance_code)

    def _search_similarity_unit(self, unit_type, vector, limit):
        return self._search_similarity_units(unit_type, [vector], limit)[0]

    # ---- simprints (index.py:1084-1469)
    def _search_simprints(self, query, limit, exact=False):
        from iscc_search_b200.simprint import SimprintMatchMulti, TypeMatchResult

        total_assets = len(self._log.assets)
        threshold = self._opts["match_threshold_simprints"]
        asset_type_results = {}  # iscc_id_body -> {sp_type: TypeMatchResult}
        for sp_type, simprint_objs in query.simprints.items():
            sp_index = self._simprint_indexes.get(sp_type)
            if sp_index is None:
                continue
            query_sp_bytes = [ic.decode_base64(s.root if hasattr(s, "root") else s) for s in simprint_objs]
            if exact:
                type_total = len(self._log.simprints.get(sp_type) or {})
                raw_matches = sp_index.search_exact(query_sp_bytes, total_assets=type_total, limit=limit * 2,
                                                    threshold=threshold, detailed=True)
            else:
                raw_matches = sp_index.search_raw(simprints=query_sp_bytes, limit=limit * 2, threshold=threshold, detailed=True,
                                                  doc_freq_fn="index", total_assets=total_assets)
            for raw in raw_matches:
                asset_type_results.setdefault(raw.iscc_id_body, {})[sp_type] = TypeMatchResult(
                    score=raw.score, queried=raw.queried, matches=raw.matches, chunks=raw.chunks)
        if not asset_type_results:
            return []
        multi = []
        for iscc_id_body, type_results in asset_type_results.items():
            asset_score = sum(tr.score for tr in type_results.values()) / len(type_results)
            multi.append(SimprintMatchMulti(iscc_id=IsccID.from_body(iscc_id_body, self._realm_id).digest, score=asset_score,
                                            types=type_results))
        multi.sort(key=lambda x: (-x.score, x.iscc_id))
        return [self._convert_simprint_match(m) for m in multi[:limit]]

    def _convert_simprint_match(self, raw_match):
        """Bytes -> strings + metadata enrichment (index.py:1101-1163)."""
        schema = entries.schema
        source, metadata = None, None
        asset_bytes = self._log.assets.get(int.from_bytes(raw_match.iscc_id[2:], "big", signed=False))
        if asset_bytes is not None:
            asset = entries.deserialize_asset(asset_bytes)
            if asset.metadata:
                source, metadata = asset.metadata.get("source"), asset.metadata
        types_converted = {}
        for sp_type, tr in raw_match.types.items():
            chunks = None
            if tr.chunks is not None:
                chunks = [schema.IsccMatchedChunk(query=ic.encode_base64(c.query), match=ic.encode_base64(c.match), score=c.score,
                                                  freq=c.freq, offset=c.offset, size=c.size, content=None) for c in tr.chunks]
            types_converted[sp_type] = schema.Types(score=tr.score, matches=tr.matches, queried=tr.queried, chunks=chunks)
        return schema.IsccChunkMatch(iscc_id="ISCC:" + ic.encode_base32(raw_match.iscc_id), score=raw_match.score,
                                     types=types_converted, source=source, metadata=metadata)

    # ---- bookkeeping (index.py:883-1056)
    def flush(self):
        with self._write_lock:
            self._log.commit()
            for index in list(self._nphd_indexes.values()) + list(self._simprint_indexes.values()):
                if index.dirty:
                    index.save()
            self._mark_snapshots()

    def close(self):
        if self._closed:
            return
        with self._write_lock:
            if self._closed:
                return
            for door in self._doors.values():
                door.close()
            self._doors.clear()
            all_saved = True
            for index in list(self._nphd_indexes.values()) + list(self._simprint_indexes.values()):
                try:
                    if index.dirty:
                        index.save()
                except Exception:  # pragma: no cover - one failing store must not keep the others from saving
                    all_saved = False
            self._log.close()  # may compact the log: the marker is written after it
            if all_saved:
                self._mark_snapshots()
            for index in list(self._nphd_indexes.values()) + list(self._simprint_indexes.values()):
                try:
                    index.close()
                except Exception:  # pragma: no cover
                    pass
            self._nphd_indexes.clear()
            self._simprint_indexes.clear()
            self._instance.close()
            self._closed = True

    def __len__(self):
        return len(self._log.assets)

    @property
    def derived_sizes(self):
        # type: () -> dict[str, int]
        sizes = {t: int(ix.serialized_length) for t, ix in self._nphd_indexes.items()}
        for sp_type, sp_index in self._simprint_indexes.items():
            sizes[f"SIMPRINT_{sp_type}"] = int(sp_index.data_size)
        return sizes


def _as_query_simprints(simprints):
    """Entry simprints ({type: [IsccSimprint]}) -> query form ({type: [base64 str]}), as IsccQuery's validation does."""
    if not simprints:
        return None
    return {t: [s.simprint if hasattr(s, "simprint") else s for s in lst] for t, lst in simprints.items()}


class B200IndexManager:
    """IsccIndexProtocol over a directory of `B200Index` sub-directories. Mirrors `UsearchIndexManager`."""

    MARKER = AssetLog.META   # (either of ASSET_STORE_MARKERS marks an index directory)

    def __init__(self, base_path, max_dim=256, device=0, stores=None, **options):
        self.base_path = Path(base_path)
        self.base_path.mkdir(parents=True, exist_ok=True)
        self.max_dim = max_dim
        self._device, self._stores, self._options = device, stores, options
        self._index_cache = {}
        self._cache_lock = threading.Lock()

    def list_indexes(self):
        schema = entries.schema
        found = []
        for index_dir in self.base_path.iterdir():
            if not index_dir.is_dir() or not any((index_dir / m).exists() for m in ASSET_STORE_MARKERS):
                continue
            try:
                idx = self._get_or_load_index(index_dir.name)
                size_mb, sizes_mb = self._get_index_sizes_mb(idx)
                found.append(schema.IsccIndex(name=index_dir.name, assets=len(idx), size=size_mb, sizes=sizes_mb))
            except Exception:  # corrupted or inaccessible index: skipped, like manager.py:83-86
                continue
        found.sort(key=lambda x: x.name)
        return found

    def create_index(self, index):
        entries.validate_index_name(index.name)
        index_path = self.base_path / index.name
        if index_path.exists():
            raise FileExistsError(f"Index '{index.name}' already exists")
        try:
            self._index_cache[index.name] = self._open(index_path)
        except Exception:
            shutil.rmtree(index_path, ignore_errors=True)  # e.g. no CUDA device: leave no half-created index behind
            raise
        return entries.schema.IsccIndex(name=index.name, assets=0, size=0)

    def get_index(self, name):
        self._validate_index_exists(name)
        idx = self._get_or_load_index(name)
        size_mb, sizes_mb = self._get_index_sizes_mb(idx)
        return entries.schema.IsccIndex(name=name, assets=len(idx), size=size_mb, sizes=sizes_mb)

    def delete_index(self, name):
        self._validate_index_exists(name)
        if name in self._index_cache:
            self._index_cache[name].close()
            del self._index_cache[name]
        shutil.rmtree(self.base_path / name)

    def add_assets(self, index_name, assets):
        self._validate_index_exists(index_name)
        return self._get_or_load_index(index_name).add_assets(assets)

    def get_asset(self, index_name, iscc_id):
        self._validate_index_exists(index_name)
        return self._get_or_load_index(index_name).get_asset(iscc_id)

    def search_assets(self, index_name, query, limit=100):
        self._validate_index_exists(index_name)
        return self._get_or_load_index(index_name).search_assets(query, limit)

    def search_assets_batch(self, index_name, queries, limit=100):
        self._validate_index_exists(index_name)
        return self._get_or_load_index(index_name).search_assets_batch(queries, limit)

    def release_scratch(self, index_name=None):
        # type: (str | None) -> int
        """Free the GPU working memory of one index, or of every index this manager has open (extension; rows stay resident)."""
        if index_name is not None:
            self._validate_index_exists(index_name)
            indexes = [self._get_or_load_index(index_name)]
        else:
            with self._cache_lock:
                indexes = list(self._index_cache.values())
        return sum(idx.release_scratch() for idx in indexes)

    def rebuild(self, name, unit_types=None, simprint_types=None):
        self._validate_index_exists(name)
        idx = self._get_or_load_index(name)
        if unit_types is None:
            unit_types = idx.tracked_unit_types
        if simprint_types is None:
            simprint_types = idx.tracked_simprint_types
        return idx.rebuild(unit_types, simprint_types)

    def close(self):
        for _name, idx in list(self._index_cache.items()):
            try:
                idx.close()
            except Exception:  # pragma: no cover
                pass
        self._index_cache = {}

    # -- helpers
    def _open(self, index_path):
        return B200Index(index_path, realm_id=None, max_dim=self.max_dim, device=self._device, stores=self._stores, **self._options)

    def _get_or_load_index(self, name):
        if name in self._index_cache:
            return self._index_cache[name]
        with self._cache_lock:
            if name not in self._index_cache:
                self._index_cache[name] = self._open(self.base_path / name)
            return self._index_cache[name]

    def _validate_index_exists(self, name):
        if not any((self.base_path / name / m).exists() for m in ASSET_STORE_MARKERS):
            raise FileNotFoundError(f"Index '{name}' not found")

    @staticmethod
    def _get_index_sizes_mb(idx):
        # type: (B200Index) -> tuple[int, dict[str, int]]
        """(total MB, per-component MB); component "lmdb" names the host log for compatibility with manager.py:296-335."""
        component_bytes = {"lmdb": idx._log.used_bytes()}
        component_bytes.update(idx.derived_sizes)
        mb = 1024 * 1024
        return sum(component_bytes.values()) // mb, {name: size // mb for name, size in component_bytes.items()}


def get_index(uri, **options):
    # type: (str, object) -> B200IndexManager
    """
    `b200:///path[?device=0,1,2,3]` -> `B200IndexManager` - the branch iscc-search's `options.get_index()`
    (options.py:327-375) needs next to `lmdb://` and `usearch://`. Options are the `DEFAULT_OPTIONS` keys.
    """
    from urllib.parse import parse_qs, urlparse

    parsed = urlparse(uri)
    if parsed.scheme != "b200":
        raise ValueError(f"Unsupported index URI scheme: '{uri}'. Expected b200:///path")
    path = parsed.path[1:] if parsed.path.startswith("//") else parsed.path
    if not path:
        raise ValueError(f"Index URI '{uri}' names no directory")
    query = parse_qs(parsed.query)
    devices = [int(d) for d in ",".join(query.get("device", ["0"])).split(",") if d != ""]
    return B200IndexManager(path, device=devices if len(devices) > 1 else devices[0], **options)
