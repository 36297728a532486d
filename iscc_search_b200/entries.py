"""
Host-side helpers of the protocol backend: entry (de)serialisation, ISCC-ID validation, query normalisation.

Behaviour contract: /root/reference/iscc_search/indexes/common.py (`serialize_asset :28-41`, `validate_iscc_id
:223-272`, `normalize_query :275-330`). The error TEXTS are part of that contract - the reference's own flow,
recorded in tests/golden/protocol_flow.json, is replayed against this backend message for message, and the REST
conformance tests pin substrings of them - so they are kept as data in `_MSG`. The code around
them is this package's own: one parser (`parse_iscc_id`) yields the fields of an ISCC-ID, a rule table decides
what is wrong with them, and every public helper is a view on the parsed value.
"""

import json
import re
from typing import NamedTuple

from iscc_search_b200 import iscc as ic
from iscc_search_b200 import schema as _default_schema
from iscc_search_b200.iscc import IsccCode, IsccUnit

INDEX_NAME_PATTERN = re.compile(r"^[a-z][a-z0-9]*$")

schema = _default_schema  # replaced by backend.set_schema() inside iscc-search

_COMPACT_JSON = json.JSONEncoder(separators=(",", ":"))

# message templates pinned by the reference's recorded flow / REST tests
_MSG = {
    "index_name": "Invalid index name: '{name}'. Must match pattern ^[a-z][a-z0-9]*$ "
                  "(start with lowercase letter, followed by lowercase letters/digits only)",
    "prefix": "Invalid ISCC-ID format: '{text}' (must start with 'ISCC:')",
    "base32": "Invalid ISCC-ID base32 encoding: {error}",
    "length": "Invalid ISCC-ID length: {n} bytes (expected 10 bytes = 2-byte header + 8-byte body)",
    "main_type": "Invalid ISCC-ID main type: {main_type} (expected {expected})",
    "length_field": "Invalid ISCC-ID length field: {length_field} (expected 0 for 64-bit ISCC-ID v1). "
                    "ISCC-ID '{text}' appears to be malformed.",
    "realm": "Realm mismatch: ISCC-ID '{text}' has realm={realm}, but expected realm={expected}. "
             "Cannot query assets from different realm.",
    "realm_id": "Invalid realm_id {realm}, must be 0 or 1",
    "body": "ISCC-ID body must be 8 bytes, got {n}",
    "empty_query": "Query must have 'iscc_code', 'units', or 'simprints' for search",
}


class IsccIdParts(NamedTuple):
    """Decoded ISCC-ID: header fields + the 8-byte body that is the uint64 key of the unit stores."""

    text: str
    main_type: int
    realm: int
    version: int
    length_field: int
    body: bytes


# (field predicate that must hold, message key) - evaluated in order on a decoded header
_HEADER_RULES = (
    (lambda p: p.main_type == ic.MT.ID, "main_type"),
    (lambda p: p.length_field == 0, "length_field"),
)


def parse_iscc_id(text, expected_realm=None):
    # type: (str, int | None) -> IsccIdParts
    """Decode and check an ISCC-ID string; raises ValueError with the reference's message for the first rule it breaks."""
    if not text or not text.startswith("ISCC:"):
        raise ValueError(_MSG["prefix"].format(text=text))
    try:
        raw = ic.decode_base32(text.rsplit(":", 1)[-1])
    except Exception as error:
        raise ValueError(_MSG["base32"].format(error=error))
    if len(raw) != 10:
        raise ValueError(_MSG["length"].format(n=len(raw)))
    main_type, realm, version, length_field, body = ic.decode_header(raw)
    parts = IsccIdParts(text, main_type, realm, version, length_field, bytes(raw[2:]))
    for holds, key in _HEADER_RULES:
        if not holds(parts):
            raise ValueError(_MSG[key].format(text=text, main_type=main_type, expected=ic.MT.ID, length_field=length_field))
    if expected_realm is not None and realm != expected_realm:
        raise ValueError(_MSG["realm"].format(text=text, realm=realm, expected=expected_realm))
    return parts


def validate_iscc_id(iscc_id, expected_realm=None):
    # type: (str, int | None) -> None
    parse_iscc_id(iscc_id, expected_realm)


def extract_iscc_id_body(iscc_id):
    # type: (str) -> bytes
    return parse_iscc_id(iscc_id).body


def extract_realm_id(iscc_id):
    # type: (str) -> int
    return parse_iscc_id(iscc_id).realm


def reconstruct_iscc_id(body, realm_id):
    # type: (bytes, int) -> str
    """Inverse of `parse_iscc_id(...).body / .realm` for ISCC-ID v1."""
    for ok, key, fields in ((realm_id in (0, 1), "realm_id", {"realm": realm_id}), (len(body) == 8, "body", {"n": len(body)})):
        if not ok:
            raise ValueError(_MSG[key].format(**fields))
    return "ISCC:" + ic.encode_base32(ic.encode_header(ic.MT.ID, realm_id, ic.VS.V1, 0) + body)


def validate_index_name(name):
    # type: (str) -> None
    if INDEX_NAME_PATTERN.match(name) is None:
        raise ValueError(_MSG["index_name"].format(name=name))


def serialize_asset(asset):
    # type: (IsccEntry) -> bytes
    """Stored form of an entry: compact JSON, unset fields dropped, simprints left out (they live in the simprint stores)."""
    payload = asset.model_dump(mode="json", exclude_none=True)
    payload.pop("simprints", None)
    return _COMPACT_JSON.encode(payload).encode("utf-8")


def deserialize_asset(data):
    # type: (bytes) -> IsccEntry
    return schema.IsccEntry(**json.loads(data))


def extract_unit_body(unit):
    # type: (str) -> bytes
    return IsccUnit(unit).body


def get_unit_type(unit):
    # type: (str) -> str
    return IsccUnit(unit).unit_type


def _with_composed_code(query):
    """units -> also the ISCC-CODE they compose, when they compose one (other unit sets stay searchable as they are)."""
    try:
        return query.model_copy(update={"iscc_code": ic.gen_iscc_code(query.units, wide=True)["iscc"]})
    except ValueError:
        return query


def _with_decomposed_units(query):
    return query.model_copy(update={"units": [str(unit) for unit in IsccCode(query.iscc_code).units]})


# (has units, has iscc_code) -> completion of the missing representation
_NORMALIZERS = {
    (True, True): lambda query: query,
    (True, False): _with_composed_code,
    (False, True): _with_decomposed_units,
}


def normalize_query(query):
    # type: (IsccQuery) -> IsccQuery
    """Give a query both representations (units and ISCC-CODE) when one is derivable from the other; simprint-only queries pass."""
    shape = (bool(query.units), bool(query.iscc_code))
    if shape in _NORMALIZERS:
        return _NORMALIZERS[shape](query)
    if query.simprints:
        return query
    raise ValueError(_MSG["empty_query"])
