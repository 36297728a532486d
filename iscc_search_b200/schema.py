"""
Result and request models of the index protocol.

Inside the reference these classes come from `iscc_search.schema` (generated from its OpenAPI document,
/root/reference/iscc_search/schema.py). The backend in `backend.py` only needs their field names and validation
rules, so this module declares the same models compactly (same class names, fields, bounds and patterns, none of
the generated documentation). When the backend is dropped into iscc-search, `backend.set_schema(iscc_search.schema)`
makes it return the reference's own classes.
"""

from enum import StrEnum
from typing import Annotated, Any

from pydantic import AnyUrl, BaseModel, ConfigDict, Field, RootModel

_ISCC_ID = Field(pattern="^ISCC:[A-Z2-7]{16}$")
_ISCC_CODE = Field(pattern="^ISCC:[A-Z2-7]{16,}$")
_B64 = Field(pattern="^[A-Za-z0-9+/_=-]+$")
U32 = Annotated[int, Field(ge=0, le=4294967295)]
UnitScore = Annotated[float, Field(ge=0.0, le=1.0)]


class HttpError(BaseModel):
    detail: str | list[str]


class IsccIndex(BaseModel):
    name: Annotated[str, Field(min_length=1, max_length=32, pattern="^[a-z][a-z0-9]*$")]
    assets: Annotated[int | None, Field(ge=0)] = None
    size: Annotated[int | None, Field(ge=0)] = None
    sizes: dict[str, Annotated[int, Field(ge=0)]] | None = None


class IsccSimprint(BaseModel):
    simprint: Annotated[str, Field(min_length=11, pattern="^[A-Za-z0-9+/_=-]+$")]
    offset: U32
    size: U32


class QuerySimprint(RootModel[str]):
    root: Annotated[str, _B64]


class IsccQuery(BaseModel):
    iscc_id: Annotated[str | None, _ISCC_ID] = None
    iscc_code: Annotated[str | None, _ISCC_CODE] = None
    units: Annotated[list[str] | None, Field(min_length=1)] = None
    simprints: dict[str, Annotated[list[QuerySimprint], Field(min_length=1)]] | None = None


class Status(StrEnum):
    created = "created"
    updated = "updated"


class IsccAddResult(BaseModel):
    iscc_id: Annotated[str, _ISCC_ID]
    status: Status


class IsccMetadata(BaseModel):
    model_config = ConfigDict(extra="allow")
    name: str | None = None
    source: AnyUrl | None = None


class IsccMatchedChunk(BaseModel):
    query: Annotated[str, _B64]
    match: Annotated[str, _B64]
    score: UnitScore
    freq: Annotated[int, Field(ge=1)]
    offset: U32
    size: U32
    content: Annotated[str | None, Field(pattern="^data:[^;]+;base64,.+$")] = None


class IsccEntry(BaseModel):
    iscc_id: Annotated[str | None, _ISCC_ID] = None
    iscc_code: Annotated[str | None, _ISCC_CODE] = None
    units: Annotated[list[str] | None, Field(min_length=2)] = None
    simprints: dict[str, Annotated[list[IsccSimprint], Field(min_length=1)]] | None = None
    metadata: dict[str, Any] | None = None


class IsccGlobalMatch(BaseModel):
    iscc_id: Annotated[str, _ISCC_ID]
    score: UnitScore
    types: Annotated[dict[str, UnitScore], Field(min_length=1)]
    metadata: IsccMetadata | None = None


class Types(BaseModel):
    score: UnitScore
    matches: Annotated[int, Field(ge=0)]
    queried: Annotated[int, Field(ge=1)]
    chunks: list[IsccMatchedChunk] | None = None


class IsccChunkMatch(BaseModel):
    iscc_id: Annotated[str, _ISCC_ID]
    score: UnitScore
    types: Annotated[dict[str, Types], Field(min_length=1)]
    source: AnyUrl | None = None
    metadata: IsccMetadata | None = None


class IsccSearchResult(BaseModel):
    query: IsccQuery
    global_matches: list[IsccGlobalMatch]
    chunk_matches: list[IsccChunkMatch] = []
