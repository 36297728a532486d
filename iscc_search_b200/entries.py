"""
Host-side helpers shared by the protocol backend: entry (de)serialisation, ISCC-ID validation, query normalisation.

Same function names, arguments, results and error messages as /root/reference/iscc_search/indexes/common.py
(the server tests pin substrings of these messages), written against this package's own codec
(`iscc.py`) and schema classes.
"""

import json
import re

from iscc_search_b200 import iscc as ic
from iscc_search_b200 import schema as _default_schema
from iscc_search_b200.iscc import IsccCode, IsccUnit

INDEX_NAME_PATTERN = re.compile(r"^[a-z][a-z0-9]*$")

schema = _default_schema  # replaced by backend.set_schema() inside iscc-search


def serialize_asset(asset):
    # type: (IsccEntry) -> bytes
    """Compact JSON without `simprints` (they live in the simprint stores) - common.py:28-41."""
    d = asset.model_dump(mode="json", exclude_none=True, exclude={"simprints"})
    return json.dumps(d, separators=(",", ":")).encode("utf-8")


def deserialize_asset(data):
    # type: (bytes) -> IsccEntry
    return schema.IsccEntry(**json.loads(data.decode("utf-8")))


def validate_index_name(name):
    # type: (str) -> None
    if not INDEX_NAME_PATTERN.match(name):
        raise ValueError(
            f"Invalid index name: '{name}'. Must match pattern ^[a-z][a-z0-9]*$ "
            f"(start with lowercase letter, followed by lowercase letters/digits only)"
        )


def validate_iscc_id(iscc_id, expected_realm=None):
    # type: (str, int | None) -> None
    """Format, length, MainType, length field and (optionally) realm of an ISCC-ID - common.py:214-267."""
    if not iscc_id or not iscc_id.startswith("ISCC:"):
        raise ValueError(f"Invalid ISCC-ID format: '{iscc_id}' (must start with 'ISCC:')")
    try:
        code_bytes = ic.decode_base32(iscc_id.split(":")[-1])
    except Exception as e:
        raise ValueError(f"Invalid ISCC-ID base32 encoding: {e}")
    if len(code_bytes) != 10:
        raise ValueError(
            f"Invalid ISCC-ID length: {len(code_bytes)} bytes (expected 10 bytes = 2-byte header + 8-byte body)"
        )
    mt, realm, _vs, length_field, _body = ic.decode_header(code_bytes)
    if mt != ic.MT.ID:
        raise ValueError(f"Invalid ISCC-ID main type: {mt} (expected {ic.MT.ID})")
    if length_field != 0:
        raise ValueError(
            f"Invalid ISCC-ID length field: {length_field} (expected 0 for 64-bit ISCC-ID v1). "
            f"ISCC-ID '{iscc_id}' appears to be malformed."
        )
    if expected_realm is not None and realm != expected_realm:
        raise ValueError(
            f"Realm mismatch: ISCC-ID '{iscc_id}' has realm={realm}, but expected realm={expected_realm}. "
            f"Cannot query assets from different realm."
        )


def extract_iscc_id_body(iscc_id):
    # type: (str) -> bytes
    validate_iscc_id(iscc_id)
    return ic.decode_base32(iscc_id.split(":")[-1])[2:]


def extract_realm_id(iscc_id):
    # type: (str) -> int
    validate_iscc_id(iscc_id)
    return ic.decode_header(ic.decode_base32(iscc_id.split(":")[-1]))[1]


def reconstruct_iscc_id(body, realm_id):
    # type: (bytes, int) -> str
    if realm_id not in (0, 1):
        raise ValueError(f"Invalid realm_id {realm_id}, must be 0 or 1")
    if len(body) != 8:
        raise ValueError(f"ISCC-ID body must be 8 bytes, got {len(body)}")
    return "ISCC:" + ic.encode_base32(ic.encode_header(ic.MT.ID, realm_id, ic.VS.V1, 0) + body)


def extract_unit_body(unit):
    # type: (str) -> bytes
    return IsccUnit(unit).body


def get_unit_type(unit):
    # type: (str) -> str
    return IsccUnit(unit).unit_type


def normalize_query(query):
    # type: (IsccQuery) -> IsccQuery
    """
    Both representations when possible (common.py:270-330): units -> try to compose the ISCC-CODE,
    ISCC-CODE -> decompose into units, simprints-only passes through, nothing -> ValueError.
    """
    if query.units and query.iscc_code:
        return query
    if query.units and not query.iscc_code:
        try:
            return query.model_copy(update={"iscc_code": ic.gen_iscc_code(query.units, wide=True)["iscc"]})
        except ValueError:
            return query  # units that do not form a valid ISCC-CODE are still searchable
    if query.iscc_code and not query.units:
        return query.model_copy(update={"units": [str(u) for u in IsccCode(query.iscc_code).units]})
    if query.simprints:
        return query
    raise ValueError("Query must have 'iscc_code', 'units', or 'simprints' for search")
