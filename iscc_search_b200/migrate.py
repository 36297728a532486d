"""
One-way import of an existing iscc-search `usearch:///` index directory into a `b200:///` one.

The reference keeps its source of truth in `index.lmdb` (/root/reference/iscc_search/indexes/usearch/index.py:87-103):

    __metadata__                 realm_id (>I), max_dim (>I), created_at (>d), sp_types (JSON list)
    __assets__                   ISCC-ID body (>Q) -> asset JSON without simprints
    __sp_{type}__      dupsort   simprint bytes -> 16-byte chunk pointer (ISCC-ID body | offset | size)   (:1195-1231)
    __sp_assets_{type}__         ISCC-ID body -> 16-byte fingerprint of the asset's entries, b"" for legacy assets
    __instance__       dupsort   INSTANCE body -> ISCC-ID body                       (derivable from the asset units)

and derived HNSW shard directories next to it. Only the LMDB tables are read: they are replayed into this package's
`AssetLog`; the HBM stores are then built from the log the first time `B200Index` opens the new directory - the same
thing the reference's own `rebuild()` does from LMDB (index.py:1058-1082, 1650-1726, 1829-1885).

    python -m iscc_search_b200.migrate /data/iscc-search/myindex /data/b200/myindex

needs the `lmdb` package (it ships with iscc-search); `import_lmdb_env` takes any object with py-lmdb's
`begin / open_db / cursor` surface.
"""

import json
import struct
import sys
from pathlib import Path

from iscc_search_b200.assetlog import AssetLog
from iscc_search_b200.backend import SP_FINGERPRINT_BYTES, simprint_fingerprint
from iscc_search_b200.simprint import unpack_chunk_pointer


class _Entry:
    """(simprint bytes, offset, size) with the attribute names `simprint_fingerprint` reads."""

    __slots__ = ("simprint", "offset", "size")

    def __init__(self, sp_bytes, offset, size):
        from iscc_search_b200 import iscc as ic

        self.simprint, self.offset, self.size = ic.encode_base64(sp_bytes), offset, size


def import_lmdb_env(env, dst_path):
    # type: (object, str | Path) -> dict
    """
    Replay the tables of an open LMDB environment of a reference index into a new `AssetLog` at `dst_path`.

    :return: counts {"assets", "simprint_types", "simprint_entries", "realm_id", "max_dim"}
    """
    dst = Path(dst_path)
    if (dst / AssetLog.META).exists():
        raise FileExistsError(f"'{dst}' already holds an index")
    with env.begin() as txn:
        meta_db = env.open_db(b"__metadata__", txn=txn)
        realm_raw, dim_raw = txn.get(b"realm_id", db=meta_db), txn.get(b"max_dim", db=meta_db)
        realm_id = struct.unpack(">I", realm_raw)[0] if realm_raw is not None else None
        max_dim = struct.unpack(">I", dim_raw)[0] if dim_raw is not None else 256
        sp_types_raw = txn.get(b"sp_types", db=meta_db)
        sp_types = json.loads(sp_types_raw.decode()) if sp_types_raw else []
        log = AssetLog(dst, realm_id=realm_id, max_dim=max_dim)
        created = txn.get(b"created_at", db=meta_db)
        if created is not None:
            log.created_at = struct.unpack(">d", created)[0]
            log._write_meta()

        n_assets = 0
        assets_db = env.open_db(b"__assets__", txn=txn)
        for key_bytes, asset_bytes in txn.cursor(assets_db):
            log.put_asset(struct.unpack(">Q", bytes(key_bytes))[0], bytes(asset_bytes))
            n_assets += 1

        n_entries = 0
        for sp_type in sp_types:
            data_db = env.open_db(f"__sp_{sp_type}__".encode(), txn=txn, dupsort=True, dupfixed=True)
            marks_db = env.open_db(f"__sp_assets_{sp_type}__".encode(), txn=txn)
            per_asset = {}  # ISCC-ID body -> [(simprint bytes, offset, size)]
            for sp_bytes, pointer in txn.cursor(data_db):
                body, offset, size = unpack_chunk_pointer(bytes(pointer))
                per_asset.setdefault(body, []).append((bytes(sp_bytes), offset, size))
            for body, sp_entries in per_asset.items():
                stored = txn.get(body, db=marks_db)
                if stored is not None and len(stored) == SP_FINGERPRINT_BYTES:
                    fingerprint = bytes(stored)
                else:  # legacy marker (index.py:612-618): the fingerprint is recomputed from the stored entries
                    fingerprint = simprint_fingerprint([_Entry(*e) for e in sp_entries])
                log.put_simprints(sp_type, body, fingerprint, sp_entries)
                n_entries += len(sp_entries)
        log.close()
    return {"assets": n_assets, "simprint_types": sp_types, "simprint_entries": n_entries, "realm_id": realm_id, "max_dim": max_dim}


def import_usearch_index(src_path, dst_path):
    # type: (str | Path, str | Path) -> dict
    """Open `src_path/index.lmdb` read-only with py-lmdb and import it (see module docstring)."""
    try:
        import lmdb
    except ImportError as e:  # pragma: no cover - lmdb is not part of this image
        raise ImportError("importing a usearch:/// index needs the `lmdb` package (a dependency of iscc-search)") from e
    lmdb_file = Path(src_path) / "index.lmdb"
    if not lmdb_file.exists():
        raise FileNotFoundError(f"'{lmdb_file}' not found")
    env = lmdb.open(str(lmdb_file), subdir=False, readonly=True, lock=False, max_dbs=64)  # pragma: no cover
    try:  # pragma: no cover
        return import_lmdb_env(env, dst_path)
    finally:  # pragma: no cover
        env.close()


if __name__ == "__main__":  # pragma: no cover
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    print(json.dumps(import_usearch_index(sys.argv[1], sys.argv[2])))
