"""
Host-side source of truth of one index: asset JSON by ISCC-ID key, per-asset simprint entries, index metadata.

The reference keeps these in LMDB tables (`__metadata__`, `__assets__`, `sp_assets:*`, `sp_data:*` -
/root/reference/iscc_search/indexes/usearch/index.py:87-103, 1165-1231). LMDB is not part of this image and is out
of scope for the GPU path ("LMDB keeps only the host-side metadata and key mapping"), so this is the
smallest durable stand-in with the same role: dictionaries in memory, mirrored to an append-only record log that is
replayed on open (last record wins). The HBM stores are derived data and can always be rebuilt from it.

    index.meta.json           {"realm_id", "max_dim", "created_at"}
    assets.log                records: type(1) | payload length(4, big-endian) | payload
        'A'  key(8) | asset JSON                        (put / overwrite)
        'S'  JSON {"t": type, "b": body hex, "f": fingerprint hex, "e": [[simprint hex, offset, size], ...]}
"""

import json
import os
import struct
import time
from pathlib import Path


class AssetLog:
    META = "index.meta.json"
    LOG = "assets.log"

    def __init__(self, path, realm_id=None, max_dim=256):
        # type: (str | Path, int | None, int) -> None
        self.path = Path(path)
        self.path.mkdir(parents=True, exist_ok=True)
        self.assets = {}      # type: dict[int, bytes]
        self.simprints = {}   # type: dict[str, dict[bytes, tuple[bytes, list[tuple[bytes, int, int]]]]]
        self.stale_records = 0
        meta_file = self.path / self.META
        if meta_file.exists():
            meta = json.loads(meta_file.read_text())
            self.realm_id, self.max_dim, self.created_at = meta["realm_id"], meta["max_dim"], meta["created_at"]
        else:
            self.realm_id, self.max_dim, self.created_at = realm_id, max_dim, time.time()
            self._write_meta()
        self._replay()
        self._fh = open(self.path / self.LOG, "ab")

    # -- metadata
    def _write_meta(self):
        tmp = self.path / (self.META + ".tmp")
        tmp.write_text(json.dumps({"realm_id": self.realm_id, "max_dim": self.max_dim, "created_at": self.created_at}))
        os.replace(tmp, self.path / self.META)

    def set_realm(self, realm_id):
        # type: (int) -> None
        self.realm_id = realm_id
        self._write_meta()

    # -- records
    def _replay(self):
        log = self.path / self.LOG
        if not log.exists():
            return
        data = log.read_bytes()
        pos, good = 0, 0
        while pos + 5 <= len(data):
            kind, n = data[pos:pos + 1], struct.unpack(">I", data[pos + 1:pos + 5])[0]
            if pos + 5 + n > len(data):
                break  # torn tail of an interrupted append
            payload = data[pos + 5:pos + 5 + n]
            if kind == b"A":
                key = struct.unpack(">Q", payload[:8])[0]
                self.stale_records += key in self.assets
                self.assets[key] = payload[8:]
            elif kind == b"S":
                d = json.loads(payload)
                table = self.simprints.setdefault(d["t"], {})
                body = bytes.fromhex(d["b"])
                self.stale_records += body in table
                table[body] = (bytes.fromhex(d["f"]), [(bytes.fromhex(s), o, z) for s, o, z in d["e"]])
            pos += 5 + n
            good = pos
        if good != len(data):
            with open(log, "r+b") as fh:
                fh.truncate(good)

    def _append(self, kind, payload):
        self._fh.write(kind + struct.pack(">I", len(payload)) + payload)

    def put_asset(self, key, asset_bytes):
        # type: (int, bytes) -> None
        self.stale_records += key in self.assets
        self.assets[key] = asset_bytes
        self._append(b"A", struct.pack(">Q", key) + asset_bytes)

    def put_simprints(self, sp_type, body, fingerprint, entries):
        # type: (str, bytes, bytes, list[tuple[bytes, int, int]]) -> None
        table = self.simprints.setdefault(sp_type, {})
        self.stale_records += body in table
        table[body] = (fingerprint, entries)
        self._append(b"S", json.dumps({"t": sp_type, "b": body.hex(), "f": fingerprint.hex(),
                                       "e": [[s.hex(), o, z] for s, o, z in entries]}, separators=(",", ":")).encode())

    def commit(self):
        self._fh.flush()

    def compact(self):
        """Rewrite the log with live records only."""
        self._fh.close()
        tmp = self.path / (self.LOG + ".tmp")
        with open(tmp, "wb") as fh:
            self._fh = fh
            for key, asset_bytes in self.assets.items():
                self._append(b"A", struct.pack(">Q", key) + asset_bytes)
            for sp_type, table in self.simprints.items():
                for body, (fingerprint, entries) in table.items():
                    self._append(b"S", json.dumps({"t": sp_type, "b": body.hex(), "f": fingerprint.hex(),
                                                   "e": [[s.hex(), o, z] for s, o, z in entries]}, separators=(",", ":")).encode())
        os.replace(tmp, self.path / self.LOG)
        self.stale_records = 0
        self._fh = open(self.path / self.LOG, "ab")

    def log_bytes(self):
        # type: () -> int
        """Size of the record log including buffered appends: the version stamp the derived-store snapshots are tied to."""
        if self._fh is not None:
            self._fh.flush()
        return (self.path / self.LOG).stat().st_size

    def used_bytes(self):
        # type: () -> int
        self._fh.flush()
        return sum(f.stat().st_size for f in (self.path / self.LOG, self.path / self.META) if f.exists())

    def close(self):
        if self._fh is None:
            return
        live = len(self.assets) + sum(len(t) for t in self.simprints.values())
        if self.stale_records > max(1024, live):
            self.compact()
        self._fh.close()
        self._fh = None
