"""
ISCC codec restatement (iscc_search_b200/iscc.py) against the literal ISCCs the reference ships: the OpenAPI examples
(/root/reference/iscc_search/openapi/IsccQuery.yaml, IsccSearchResult.yaml, IsccEntry.yaml -> schema.py:70-98, 350)
and the type names its tests assert (tests/test_models_iscc_id.py:106-110).
"""

import pytest

from iscc_search_b200 import entries
from iscc_search_b200 import iscc as ic
from iscc_search_b200.schema import IsccQuery

# IsccQuery example: one ISCC-CODE and the five units it is made of
CODE = "ISCC:KADUHBUDQUT3LPWRJH6BUAG7HMBIXX6JRQRX3JH7EBIOSMXEVL5URBBUPOIOTU4HLSSQ"
UNITS = ["ISCC:AAAUHBUDQUT3LPWR", "ISCC:CAAUT7A2ADPTWAUL", "ISCC:EAA57SMMEN62J7ZA", "ISCC:GAAVB2JS4SVPWSEE", "ISCC:IAATI64Q5HJYOXFF"]


def test_code_decomposes_into_the_documented_units_and_back():
    code = ic.IsccCode(CODE)
    assert code.iscc_type == "ISCC_TEXT_V0" and len(code) == 5 * 64
    assert [str(u) for u in code.units] == UNITS
    assert [u.unit_type for u in code.units] == ["META_NONE_V0", "SEMANTIC_TEXT_V0", "CONTENT_TEXT_V0", "DATA_NONE_V0", "INSTANCE_NONE_V0"]
    assert ic.gen_iscc_code(UNITS, wide=True)["iscc"] == CODE
    assert ic.gen_iscc_code(list(reversed(UNITS)))["iscc"] == CODE  # units are sorted by MainType
    assert str(code) == CODE and bytes(code) == ic.decode_base32(CODE[5:]) and ic.IsccCode(CODE[5:]).digest == code.digest


def test_long_units_are_truncated_to_64_bits_in_a_code():
    # IsccSearchResult example: 256-bit META and CONTENT-IMAGE units next to the ISCC-CODE built from their first 64 bits
    meta = ic.IsccUnit("ISCC:AADYCMZIOY36XXGZ5B5BME7EIPPXRFKYQZ7VXKI7V55AEQQE67A33BY")
    content = ic.IsccUnit("ISCC:EED7ZPIEYNACCLXXZSS2LIM6JVXDYGCG2QSMC7DCPER4MYJPJATIM4Y")
    code = ic.IsccCode("ISCC:KECYCMZIOY36XXGZ7S6QJQ2AEEXPOVEHZYPK6GMSFLU3WF54UPZMTPY")
    assert (meta.unit_type, len(meta), content.unit_type, len(content)) == ("META_NONE_V0", 256, "CONTENT_IMAGE_V0", 256)
    assert code.iscc_type == "ISCC_IMAGE_V0"
    units = code.units
    assert [u.unit_type for u in units] == ["META_NONE_V0", "CONTENT_IMAGE_V0", "DATA_NONE_V0", "INSTANCE_NONE_V0"]
    assert units[0].body == meta.body[:8] and units[1].body == content.body[:8]
    rebuilt = ic.gen_iscc_code([str(meta), str(content), str(units[2]), str(units[3])])["iscc"]
    assert rebuilt == str(code)


def test_wide_code_roundtrip_and_composition_errors():
    data = "ISCC:" + ic.encode_base32(ic.encode_header(ic.MT.DATA, 0, 0, ic.encode_length(ic.MT.DATA, 128)) + bytes(range(16)))
    inst = "ISCC:" + ic.encode_base32(ic.encode_header(ic.MT.INSTANCE, 0, 0, ic.encode_length(ic.MT.INSTANCE, 256)) + bytes(range(32, 64)))
    wide = ic.IsccCode(ic.gen_iscc_code([data, inst], wide=True)["iscc"])
    assert wide.iscc_type == "ISCC_WIDE_V0" and [len(u) for u in wide.units] == [128, 128]
    assert wide.units[0].body == bytes(range(16)) and wide.units[1].body == bytes(range(32, 48))
    narrow = ic.IsccCode(ic.gen_iscc_code([data, inst])["iscc"])
    assert narrow.iscc_type == "ISCC_SUM_V0" and [len(u) for u in narrow.units] == [64, 64]
    with pytest.raises(ValueError, match="Minimum two"):
        ic.gen_iscc_code([data])
    with pytest.raises(ValueError, match="MT.DATA and MT.INSTANCE"):
        ic.gen_iscc_code(UNITS[:3])
    with pytest.raises(ValueError, match="same SubType"):
        img = "ISCC:" + ic.encode_base32(ic.encode_header(ic.MT.CONTENT, ic.ST_CC.IMAGE, 0, 1) + bytes(8))
        ic.gen_iscc_code([UNITS[1], img, UNITS[3], UNITS[4]])


def test_iscc_id_forms():
    i = ic.IsccID("ISCC:MAIGIIFJRDGEQQAA")                 # the ISCC-ID every OpenAPI example uses
    assert i.iscc_type == "ID_REALM_0_V1" and i.realm_id == 0 and len(i) == 64 and bytes(i)[:2] == bytes([0x60, 0x10])
    assert str(ic.IsccID.from_int(int(i), 0)) == str(i) == str(ic.IsccID.from_body(i.body, 0))
    r1 = ic.IsccID.from_int(int(i), 1)
    assert r1.iscc_type == "ID_REALM_1_V1" and int(r1) == int(i) and str(r1) != str(i)
    made = ic.gen_iscc_id(timestamp=1_000_003, hub_id=7, realm_id=1)["iscc"]
    assert int(ic.IsccID(made)) == (1_000_003 << 12) | 7 and entries.extract_realm_id(made) == 1
    assert entries.reconstruct_iscc_id(entries.extract_iscc_id_body(made), 1) == made
    assert ic.IsccID.random().iscc_type == "ID_REALM_0_V1"
    assert ic.split_iscc_sequence(bytes(i) + ic.IsccUnit(UNITS[0]).digest + ic.IsccUnit(UNITS[2]).digest) == [
        bytes(i), ic.IsccUnit(UNITS[0]).digest, ic.IsccUnit(UNITS[2]).digest]
    with pytest.raises(TypeError):
        ic.IsccBase(12)


@pytest.mark.parametrize("bad, fragment", [
    ("", "must start with 'ISCC:'"), ("INVALID:ABCD1234", "must start with 'ISCC:'"), ("ISCC:INVALIDBASE32!@#", "base32"),
    ("ISCC:AAAUHBUDQUT3LPWR", "main type"), ("ISCC:MAIGIIFJRDGEQQAAAA", "length"),
])
def test_validate_iscc_id_messages(bad, fragment):
    with pytest.raises(ValueError, match=fragment):
        entries.validate_iscc_id(bad)


def test_header_varnibbles_and_text_forms():
    for values in ((0, 0, 0, 0), (7, 7, 1, 7), (2, 4, 0, 8), (5, 71, 0, 3), (6, 1, 1, 72), (0, 583, 0, 584), (1, 2, 4679, 0)):
        head = ic.encode_header(*values)
        assert ic.decode_header(head + b"\xab\xcd")[:4] == values and ic.decode_header(head + b"\xab\xcd")[4] == b"\xab\xcd"
    with pytest.raises(ValueError):
        ic.encode_header(0, 0, 0, 4680)
    for n in (8, 16, 24, 32):
        assert ic.decode_length(ic.MT.DATA, ic.encode_length(ic.MT.DATA, n * 8)) == n * 8
    assert ic.decode_base32("aaauhbudqut3lpwr") == ic.decode_base32("AAAUHBUDQUT3LPWR")
    raw = bytes(range(1, 17))
    assert ic.decode_base64(ic.encode_base64(raw)) == raw and "=" not in ic.encode_base64(raw)
    assert ic.decode_base64(ic.encode_base64(raw) + "==") == raw


def test_normalize_query_fills_the_missing_representation():
    q = entries.normalize_query(IsccQuery(iscc_code=CODE))
    assert q.units == UNITS and q.iscc_code == CODE
    q = entries.normalize_query(IsccQuery(units=UNITS))
    assert q.iscc_code == CODE
    q = entries.normalize_query(IsccQuery(units=UNITS[:2]))          # no DATA/INSTANCE: still searchable, no code
    assert q.iscc_code is None and q.units == UNITS[:2]
    sp = IsccQuery(simprints={"CONTENT_TEXT_V0": ["AXvu3tp2kF8mN9qL4rT1sZ"]})
    assert entries.normalize_query(sp) is sp
    with pytest.raises(ValueError, match="Query must have"):
        entries.normalize_query(IsccQuery())


def test_idp_example_code_matches_its_independent_datahash():
    # tests/conftest.py:295-296 of the reference: an ISCC-CODE and the multihash (0x1e20 = BLAKE3-256) of the same file.
    # The Instance-Code is the head of that hash, so the last unit of the decomposed code must equal its first 64 bits.
    code = ic.IsccCode("ISCC:KACWN77F73NA44D6EUG3S3QNJIL2BPPQFMW6ZX6CZNOKPAK23S2IJ2I")
    datahash = bytes.fromhex("1e205ca7815adcb484e9a136c11efe69c1d530176d549b5d18d038eb5280b4b3470c")
    units = code.units
    assert [u.unit_type for u in units] == ["META_NONE_V0", "CONTENT_TEXT_V0", "DATA_NONE_V0", "INSTANCE_NONE_V0"]
    assert units[-1].body == datahash[2:10]
    # the 256-bit INSTANCE unit the aggregator builds from the datahash (aggregator/entry.py:90-92) extends that unit
    full = ic.IsccUnit("ISCC:" + ic.encode_component(ic.MT.INSTANCE, ic.ST.NONE, ic.VS.V0, 256, datahash[2:]))
    assert full.unit_type == "INSTANCE_NONE_V0" and len(full) == 256 and full.body[:8] == units[-1].body
    assert ic.gen_iscc_code([str(u) for u in units])["iscc"] == str(code)
