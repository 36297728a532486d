"""
REST-level differential test: the reference's own FastAPI application (iscc_search/server, loaded unmodified from
/root/reference) serves the same request script twice - once over its stock `UsearchIndexManager`, once over this
package's `B200IndexManager` plugged in the way INTEGRATION.md section 0 describes (`backend.set_schema` + the
`get_index()` branch) - and status codes and JSON bodies must agree. Runs only where the reference tree exists.
"""

import importlib
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference")
GEN = Path(__file__).parent / "golden" / "make_protocol_golden.py"

pytestmark = pytest.mark.skipif(not (REF / "iscc_search").is_dir(), reason="the reference tree only exists in the build container")


def _load():
    from tests import test_protocol_differential as d

    gen, (index_mod, schema) = d._generator()
    pkg = sys.modules["iscc_search"]
    pkg.__version__ = "0.0.0+differential"
    ic = sys.modules["iscc_core"]
    from iscc_search_b200 import iscc as codec

    ic.encode_component = codec.encode_component
    server = importlib.import_module("iscc_search.server")
    manager_mod = importlib.import_module("iscc_search.indexes.usearch.manager")
    return gen, schema, server, manager_mod


def _script(gen):
    """[(method, path, kwargs)] - built once so both backends see identical requests."""
    from iscc_search_b200 import iscc as codec

    rng = np.random.default_rng(515)
    assets, fam = gen.make_assets(rng, 40)
    MT, ST_CC = codec.MT, codec.ST_CC
    u = gen.unit
    missing = codec.gen_iscc_id(timestamp=77, hub_id=7, realm_id=0)["iscc"]
    other_realm = codec.gen_iscc_id(timestamp=78, hub_id=7, realm_id=1)["iscc"]
    code = codec.gen_iscc_code([u(MT.META, 0, fam[0]["meta"][:8]), u(MT.CONTENT, ST_CC.TEXT, fam[0]["content"][:8]),
                                u(MT.DATA, 0, fam[0]["data"][:8]), u(MT.INSTANCE, 0, fam[0]["inst"][:8])])["iscc"]
    s = [("GET", "/indexes", {}),
         ("POST", "/indexes", {"json": {"name": "main"}}),
         ("POST", "/indexes", {"json": {"name": "main"}}),                       # 409 already exists
         ("POST", "/indexes", {"json": {"name": "Not-Valid"}}),                  # 422 schema validation
         ("POST", "/indexes", {"json": {"name": "second"}}),
         ("GET", "/indexes/ghost", {}),                                          # 404
         ("POST", "/indexes/ghost/assets", {"json": assets[:2]}),                # 404
         ("POST", "/indexes/main/assets", {"json": assets[:25]}),
         ("POST", "/indexes/main/assets", {"json": assets[20:40] + [dict(assets[3], metadata={"name": "changed"})]}),
         ("POST", "/indexes/main/assets", {"json": [{"units": assets[0]["units"]}]}),                              # 400 no iscc_id
         ("POST", "/indexes/main/assets", {"json": [dict(assets[1], iscc_id=other_realm)]}),                       # 400 realm
         ("GET", "/indexes", {}),
         ("GET", "/indexes/main", {}),
         ("GET", f"/indexes/main/assets/{assets[3]['iscc_id']}", {}),
         ("GET", f"/indexes/main/assets/{missing}", {}),                         # 404 naming the id
         ("GET", f"/indexes/main/assets/{other_realm}", {}),                     # 400 realm mismatch
         ("GET", "/indexes/main/assets/ISCC:NOTANID", {}),                       # 400 invalid id
         ("GET", f"/indexes/ghost/assets/{missing}", {})]                        # 404 index
    for f in fam[:6]:
        q = {"units": [u(MT.META, 0, f["meta"][:8]), u(MT.CONTENT, ST_CC.TEXT, f["content"]), u(MT.DATA, 0, f["data"][:16]), u(MT.INSTANCE, 0, f["inst"])]}
        s.append(("POST", "/indexes/main/search", {"json": q}))
        s.append(("POST", "/indexes/main/search", {"json": q, "params": {"limit": 3}}))
        s.append(("POST", "/indexes/main/search", {"json": {"simprints": {"CONTENT_TEXT_V0": [gen.b64(x) for x in f["sp64"]],
                                                                           "SEMANTIC_TEXT_V0": [gen.b64(x) for x in f["sp128"][:2]]}}}))
    s += [("GET", "/indexes/main/search", {"params": {"iscc_code": code}}),
          ("GET", "/indexes/main/search", {"params": {"iscc_code": code, "limit": 2}}),
          ("POST", "/indexes/main/search", {"json": {"iscc_id": assets[5]["iscc_id"]}}),
          ("POST", "/indexes/main/search", {"json": {"iscc_id": missing}}),      # 404 naming the id
          ("POST", "/indexes/main/search", {"json": {}}),                        # 400 nothing to search
          ("POST", "/indexes/ghost/search", {"json": {"iscc_code": code}}),      # 404 index
          ("POST", "/indexes/second/search", {"json": {"iscc_code": code}}),     # empty index
          ("DELETE", "/indexes/second", {}),
          ("DELETE", "/indexes/second", {}),                                     # 404
          ("GET", "/indexes", {})]
    return s


def _serve(server, manager, script, monkeypatch):
    from fastapi.testclient import TestClient

    monkeypatch.setattr(server, "get_index", lambda: manager)
    out = []
    with TestClient(server.app) as client:
        for method, path, kw in script:
            r = client.request(method, path, **kw)
            body = r.json() if r.content and r.headers.get("content-type", "").startswith("application/json") else r.text
            out.append((r.status_code, body))
    return out


def _normalise(path, body):
    """Index listings: sizes are backend-specific accounting (both floor to whole MB); everything else is compared as is."""
    if isinstance(body, list) and body and isinstance(body[0], dict) and "name" in body[0] and "assets" in body[0]:
        return [{"name": b["name"], "assets": b["assets"]} for b in body]
    if isinstance(body, dict) and set(body) >= {"name", "assets"} and "global_matches" not in body:
        return {"name": body["name"], "assets": body["assets"]}
    return body


def test_reference_rest_api_over_both_backends(tmp_path, cpu_stores, monkeypatch):
    from iscc_search_b200 import backend, entries
    from iscc_search_b200 import schema as own_schema
    from tests.protocol_replay import _same_global

    gen, schema, server, manager_mod = _load()
    script = _script(gen)
    expected = _serve(server, manager_mod.UsearchIndexManager(tmp_path / "reference"), script, monkeypatch)

    backend.set_schema(schema)  # hand the reference's own pydantic classes to the backend
    try:
        got = _serve(server, backend.B200IndexManager(tmp_path / "b200"), script, monkeypatch)
    finally:
        entries.schema = own_schema
    statuses = [e[0] for e in expected]
    assert {200, 201, 204, 400, 404, 409, 422} <= set(statuses), statuses
    for n, ((method, path, kw), (es, eb), (gs, gb)) in enumerate(zip(script, expected, got)):
        where = f"request {n}: {method} {path} {json.dumps(kw)[:120]}"
        assert gs == es, f"{where}: status {gs} != {es} ({gb})"
        if isinstance(eb, dict) and "global_matches" in eb:
            limit = int(kw.get("params", {}).get("limit", 100))
            assert gb["query"] == eb["query"], where
            _same_global(gb["global_matches"], eb["global_matches"], limit)
            assert gb.get("chunk_matches") == eb.get("chunk_matches"), where
        else:
            assert _normalise(path, gb) == _normalise(path, eb), where
