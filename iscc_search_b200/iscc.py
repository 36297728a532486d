"""
ISCC codec and value types used on either side of the search path.

The reference parses ISCC strings with the third-party `iscc_core` package (not under /root/reference and not
installable here) through its own wrappers in /root/reference/iscc_search/models.py:68-313. This module restates
the small part of the ISO 24138 codec those wrappers touch - base32/base64 text forms, the varnibble header,
length/unit-combination fields, ISCC-CODE composition and decomposition - and mirrors the wrappers
(`IsccBase`, `IsccID`, `IsccUnit`, `IsccCode`, `split_iscc_sequence`, `new_iscc_id`) with the same names,
arguments and error behaviour. Pinned by the literal ISCCs of the reference's OpenAPI examples
(tests/test_iscc_codec.py).
"""

import base64
import math
import time
from enum import IntEnum
from functools import cached_property
from random import randint

import numpy as np


class MT(IntEnum):
    META = 0
    SEMANTIC = 1
    CONTENT = 2
    DATA = 3
    INSTANCE = 4
    ISCC = 5
    ID = 6
    FLAKE = 7


class ST(IntEnum):
    NONE = 0


class ST_CC(IntEnum):
    TEXT = 0
    IMAGE = 1
    AUDIO = 2
    VIDEO = 3
    MIXED = 4


class ST_ISCC(IntEnum):
    TEXT = 0
    IMAGE = 1
    AUDIO = 2
    VIDEO = 3
    MIXED = 4
    SUM = 5
    NONE = 6
    WIDE = 7


class ST_ID(IntEnum):
    PRIVATE = 0
    BITCOIN = 1
    ETHEREUM = 2
    POLYGON = 3


class ST_ID_REALM(IntEnum):
    REALM_0 = 0
    REALM_1 = 1


class VS(IntEnum):
    V0 = 0
    V1 = 1


SUBTYPE_MAP = {
    (MT.META, VS.V0): ST,
    (MT.SEMANTIC, VS.V0): ST_CC,
    (MT.CONTENT, VS.V0): ST_CC,
    (MT.DATA, VS.V0): ST,
    (MT.INSTANCE, VS.V0): ST,
    (MT.ISCC, VS.V0): ST_ISCC,
    (MT.ID, VS.V0): ST_ID,
    (MT.ID, VS.V1): ST_ID_REALM,
    (MT.FLAKE, VS.V0): ST,
}

# optional dynamic units of an ISCC-CODE, indexed by the header's length field (bit 2 META, bit 1 SEMANTIC, bit 0 CONTENT)
UNITS = (
    (),
    (MT.CONTENT,),
    (MT.SEMANTIC,),
    (MT.SEMANTIC, MT.CONTENT),
    (MT.META,),
    (MT.META, MT.CONTENT),
    (MT.META, MT.SEMANTIC),
    (MT.META, MT.SEMANTIC, MT.CONTENT),
)

_FIXED_LENGTH_TYPES = (MT.META, MT.SEMANTIC, MT.CONTENT, MT.DATA, MT.INSTANCE, MT.FLAKE)


# ---- text forms --------------------------------------------------------------------------------
def encode_base32(data):
    # type: (bytes) -> str
    """RFC 4648 base32, upper case, no padding."""
    return base64.b32encode(data).decode("ascii").rstrip("=")


def decode_base32(code):
    # type: (str) -> bytes
    """Base32 without padding, case-insensitive."""
    pad = math.ceil(len(code) / 8) * 8 - len(code)
    return bytes(base64.b32decode(code + "=" * pad, casefold=True))


def encode_base64(data):
    # type: (bytes) -> str
    """URL-safe base64 without padding (simprints)."""
    return base64.urlsafe_b64encode(data).decode("ascii").rstrip("=")


def decode_base64(code):
    # type: (str) -> bytes
    """URL-safe base64, padding optional."""
    data = code.encode("ascii")
    return base64.urlsafe_b64decode(data + b"=" * (4 - (len(data) % 4)))


def iscc_clean(iscc):
    # type: (str) -> str
    """Strip the `ISCC:` scheme and dashes from a canonical string."""
    parts = [p.strip() for p in iscc.strip().split(":")]
    if len(parts) == 1:
        code = parts[0]
    elif len(parts) == 2:
        if parts[0].lower() != "iscc":
            raise ValueError(f"Invalid scheme: {parts[0]}")
        code = parts[1]
    else:
        raise ValueError(f"Malformed ISCC string: {iscc}")
    return code if code.startswith("u") else code.replace("-", "")


# ---- header ------------------------------------------------------------------------------------
def _varnibble_bits(n):
    # type: (int) -> str
    """Variable-length nibble code: 0xxx | 10xxxxxx | 110x(9) | 1110x(12)."""
    if 0 <= n < 8:
        return format(n, "04b")
    if 8 <= n < 72:
        return "10" + format(n - 8, "06b")
    if 72 <= n < 584:
        return "110" + format(n - 72, "09b")
    if 584 <= n < 4680:
        return "1110" + format(n - 584, "012b")
    raise ValueError("Value must be between 0 and 4679")


def _read_varnibble(bits, pos):
    # type: (str, int) -> tuple[int, int]
    left = len(bits) - pos
    if left >= 4 and bits[pos] == "0":
        return int(bits[pos:pos + 4], 2), pos + 4
    if left >= 8 and bits[pos:pos + 2] == "10":
        return int(bits[pos + 2:pos + 8], 2) + 8, pos + 8
    if left >= 12 and bits[pos:pos + 3] == "110":
        return int(bits[pos + 3:pos + 12], 2) + 72, pos + 12
    if left >= 16 and bits[pos:pos + 4] == "1110":
        return int(bits[pos + 4:pos + 16], 2) + 584, pos + 16
    raise ValueError("Invalid bitarray")


def encode_header(mtype, stype, version=0, length=1):
    # type: (int, int, int, int) -> bytes
    """MainType, SubType, Version, Length as four varnibbles, zero-padded to a byte boundary."""
    bits = "".join(_varnibble_bits(int(v)) for v in (mtype, stype, version, length))
    if len(bits) % 8:
        bits += "0000"
    return int(bits, 2).to_bytes(len(bits) // 8, "big")


def decode_header(data):
    # type: (bytes) -> tuple[int, int, int, int, bytes]
    """(MainType, SubType, Version, Length, tail bytes)."""
    head = bytes(data[:8])  # four varnibbles never need more than 8 bytes
    bits = "".join(format(b, "08b") for b in head)
    pos, out = 0, []
    for _ in range(4):
        value, pos = _read_varnibble(bits, pos)
        out.append(value)
    if pos % 8:  # strip the 4-bit padding
        if bits[pos:pos + 4] == "0000":
            pos += 4
        else:  # not byte-aligned and no padding: the tail cannot be taken byte-wise
            raise ValueError("Invalid ISCC header padding")
    out.append(bytes(data[pos // 8:]))
    return tuple(out)


def encode_length(mtype, length):
    # type: (int, int) -> int
    """Bit length (or unit combination for ISCC-CODEs) -> header length field."""
    if mtype in _FIXED_LENGTH_TYPES:
        if length >= 32 and not length % 32:
            return (length // 32) - 1
        raise ValueError(f"Invalid length {length} for MainType {mtype}")
    if mtype == MT.ISCC:
        if 0 <= length <= 7:
            return length
        raise ValueError(f"Invalid length {length} for MainType {mtype}")
    if mtype == MT.ID:
        if 64 <= length <= 96:
            return (length - 64) // 8
        raise ValueError(f"Invalid length {length} for MainType {mtype}")
    raise ValueError(f"Invalid length {length} for MainType {mtype}")


def decode_length(mtype, length, subtype=None):
    # type: (int, int, int | None) -> int
    """Header length field -> body bit length."""
    if mtype in _FIXED_LENGTH_TYPES:
        return (length + 1) * 32
    if mtype == MT.ISCC:
        if subtype == ST_ISCC.WIDE:
            return 256
        return len(decode_units(length)) * 64 + 128
    if mtype == MT.ID:
        return length * 8 + 64
    raise ValueError(f"Invalid length {length} for MainType {mtype}")


def encode_units(units):
    # type: (tuple) -> int
    return UNITS.index(tuple(units))


def decode_units(unit_id):
    # type: (int) -> tuple
    return UNITS[unit_id]


def encode_component(mtype, stype, version, bit_length, digest):
    # type: (int, int, int, int, bytes) -> str
    """Header + the first `bit_length` bits of `digest` as base32 without the `ISCC:` prefix (aggregator/entry.py:90)."""
    if mtype in (MT.ISCC, MT.ID):
        raise ValueError(f"{MT(mtype).name} is not a unit")
    nbytes = bit_length // 8
    return encode_base32(encode_header(mtype, stype, version, encode_length(mtype, bit_length)) + bytes(digest[:nbytes]))


def gen_iscc_code(codes, wide=False):
    # type: (list[str], bool) -> dict
    """
    Compose an ISCC-CODE from ISCC-UNITs (`ic.gen_iscc_code_v0`, used by the reference at
    indexes/common.py:306 and models.py:381). DATA and INSTANCE are mandatory; units are truncated to 64 bits
    (128 bits for the WIDE DATA+INSTANCE form).
    """
    codes = [iscc_clean(c) for c in codes]
    if len(codes) < 2:
        raise ValueError("Minimum two ISCC units required to generate valid ISCC-CODE")
    for c in codes:
        if len(c) < 16:
            raise ValueError(f"Cannot build ISCC-CODE from units shorter than 64-bits: {c}")
    decoded = sorted((decode_header(decode_base32(c)) for c in codes), key=lambda t: t[0])
    main_types = tuple(d[0] for d in decoded)
    if main_types[-2] != MT.DATA or main_types[-1] != MT.INSTANCE:
        raise ValueError("ISCC-CODE requires at least MT.DATA and MT.INSTANCE units.")
    is_wide = (wide and len(codes) == 2 and main_types == (MT.DATA, MT.INSTANCE)
               and all(decode_length(t[0], t[3]) >= 128 for t in decoded))
    if is_wide:
        st = ST_ISCC.WIDE
    else:
        sub_types = [t[1] for t in decoded if t[0] in (MT.SEMANTIC, MT.CONTENT)]
        if len(set(sub_types)) > 1:
            raise ValueError("Semantic-Code and Content-Code must be of same SubType")
        st = sub_types.pop() if sub_types else (ST_ISCC.SUM if len(codes) == 2 else ST_ISCC.NONE)
    encoded_length = encode_units(main_types[:-2])
    nbytes = 16 if is_wide else 8
    digest = b"".join(t[-1][:nbytes] for t in decoded)
    return {"iscc": "ISCC:" + encode_base32(encode_header(MT.ISCC, st, VS.V0, encoded_length) + digest)}


def gen_iscc_id(timestamp=None, hub_id=0, realm_id=0):
    # type: (int | None, int, int) -> dict
    """ISCC-ID v1: 52-bit microsecond timestamp + 12-bit hub id behind an ID/realm/V1 header."""
    if timestamp is None:
        timestamp = time.time_ns() // 1000
    if not 0 <= hub_id <= 4095:
        raise ValueError("hub_id must be in 0..4095")
    if timestamp >= 1 << 52:
        raise ValueError("timestamp overflow")
    body = ((timestamp << 12) | hub_id).to_bytes(8, "big")
    return {"iscc": "ISCC:" + encode_base32(encode_header(MT.ID, realm_id, VS.V1, 0) + body)}


# ---- value types (mirror of /root/reference/iscc_search/models.py) -----------------------------
def new_iscc_id():
    # type: () -> bytes
    """Random REALM-0 ISCC-ID digest (models.py:28-42)."""
    identifier = ((time.time_ns() // 1000) << 12) | randint(0, 4095)
    return encode_header(MT.ID, ST_ID_REALM.REALM_0, VS.V1, 0) + identifier.to_bytes(8, "big")


def split_iscc_sequence(data):
    # type: (bytes) -> list[bytes]
    """Split concatenated ISCC-DIGESTS (models.py:45-65)."""
    units, offset = [], 0
    try:
        while offset < len(data):
            mt, _st, _vs, ln, _body = decode_header(data[offset:])
            unit_len = 2 + decode_length(mt, ln) // 8
            units.append(data[offset:offset + unit_len])
            offset += unit_len
    except Exception as e:
        raise ValueError(f"Invalid ISCC-SEQUENCE: {e}")
    return units


class IsccBase:
    """Common representation conversions (models.py:68-153)."""

    def __init__(self, iscc):
        # type: (str | bytes) -> None
        if isinstance(iscc, str):
            self.digest = decode_base32(iscc.removeprefix("ISCC:"))
        elif isinstance(iscc, bytes):
            self.digest = iscc
        else:
            raise TypeError("`iscc` must be str, bytes")

    @property
    def body(self):
        # type: () -> bytes
        return self.digest[2:]

    @cached_property
    def fields(self):
        return decode_header(self.digest)

    @cached_property
    def iscc_type(self):
        # type: () -> str
        mt, st, vs = self.fields[0], self.fields[1], self.fields[2]
        return f"{MT(mt).name}_{SUBTYPE_MAP[(mt, vs)](st).name}_{VS(vs).name}"

    def __str__(self):
        return f"ISCC:{encode_base32(self.digest)}"

    def __len__(self):
        return len(self.digest[2:]) * 8

    def __bytes__(self):
        return self.digest


class IsccID(IsccBase):
    """ISCC-ID: header + 52-bit timestamp + 12-bit hub id (models.py:156-233)."""

    _iscc_id_headers = (encode_header(MT.ID, 0, VS.V1, 0), encode_header(MT.ID, 1, VS.V1, 0))

    def __int__(self):
        return int.from_bytes(self.body, "big", signed=False)

    @property
    def realm_id(self):
        return self.fields[1]

    @classmethod
    def from_int(cls, iscc_id, realm_id):
        return cls(cls._iscc_id_headers[realm_id] + int(iscc_id).to_bytes(8, "big", signed=False))

    @classmethod
    def from_body(cls, body, realm_id):
        return cls(cls._iscc_id_headers[realm_id] + body)

    @classmethod
    def random(cls):
        return cls(new_iscc_id())


class IsccUnit(IsccBase):
    """Single-algorithm ISCC component (models.py:236-266)."""

    @property
    def unit_type(self):
        return self.iscc_type

    def __array__(self, dtype=np.uint8, copy=None):
        arr = np.frombuffer(self.body, dtype=dtype)
        return arr.copy() if copy else arr


class IsccCode(IsccBase):
    """Composite ISCC; `units` decomposes it (models.py:269-326)."""

    @cached_property
    def units(self):
        # type: () -> list[IsccUnit]
        units, raw = [], self.digest
        while raw:
            mt, st, vs, ln, body = decode_header(raw)
            if mt != MT.ISCC:  # plain unit followed by more data
                nbytes = decode_length(mt, ln) // 8
                units.append(IsccUnit(encode_header(mt, st, vs, ln) + body[:nbytes]))
                raw = body[nbytes:]
                continue
            if st == ST_ISCC.WIDE:  # 128-bit DATA + 128-bit INSTANCE
                units.append(IsccUnit(encode_header(MT.DATA, ST.NONE, vs, encode_length(MT.DATA, 128)) + body[:16]))
                units.append(IsccUnit(encode_header(MT.INSTANCE, ST.NONE, vs, encode_length(MT.INSTANCE, 128)) + body[16:32]))
                break
            for idx, mtype in enumerate(decode_units(ln)):  # dynamic 64-bit units
                stype = ST.NONE if mtype == MT.META else st
                units.append(IsccUnit(encode_header(mtype, stype, vs, encode_length(mtype, 64)) + body[idx * 8:(idx + 1) * 8]))
            units.append(IsccUnit(encode_header(MT.DATA, ST.NONE, vs, encode_length(MT.DATA, 64)) + body[-16:-8]))
            units.append(IsccUnit(encode_header(MT.INSTANCE, ST.NONE, vs, encode_length(MT.INSTANCE, 64)) + body[-8:]))
            break
        return units
