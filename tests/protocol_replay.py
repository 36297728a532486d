"""Replays tests/golden/protocol_flow.json (steps recorded from the reference's own UsearchIndex) against `B200Index`."""

import json
from pathlib import Path

import pytest

from iscc_search_b200 import schema
from iscc_search_b200.backend import B200Index

GOLDEN = Path(__file__).parent / "golden" / "protocol_flow.json"
ERRORS = {"ValueError": ValueError, "FileNotFoundError": FileNotFoundError, "FileExistsError": FileExistsError}


def _dump(res):
    d = res.model_dump(mode="json", exclude_none=True)
    return {"query": d["query"], "global_matches": d.get("global_matches", []), "chunk_matches": d.get("chunk_matches", [])}


def _same_global(got, exp, limit):
    """
    Same matches with identical scores and per-type breakdowns. Members of a group of EQUAL total scores may come in
    another order (the reference's tie order is dict insertion order driven by LMDB cursor / usearch order, ours is
    deterministic by key), so where a tie group is cut by `limit`, only the scores of that group are compared.
    """
    assert [m["score"] for m in got] == [m["score"] for m in exp]
    assert all(a["score"] >= b["score"] for a, b in zip(got, got[1:]))
    cut_score = exp[-1]["score"] if len(exp) == limit else None
    by_id = lambda ms: sorted((m for m in ms if m["score"] != cut_score), key=lambda m: m["iscc_id"])  # noqa: E731
    assert by_id(got) == by_id(exp)


def replay(tmp_path, steps=None, **index_kwargs):
    if steps is None:
        steps = json.loads(GOLDEN.read_text())["steps"]
    idx = B200Index(tmp_path / "flow", realm_id=None, max_dim=256, **index_kwargs)
    counts = {}
    for n, step in enumerate(steps):
        op, args = step["op"], step["args"]
        counts[op] = counts.get(op, 0) + 1
        where = f"step {n} ({op})"

        def call():
            if op == "add_assets":
                return [r.model_dump(mode="json") for r in idx.add_assets([schema.IsccEntry(**a) for a in args["assets"]])]
            if op == "get_asset":
                return idx.get_asset(args["iscc_id"]).model_dump(mode="json", exclude_none=True)
            if op == "search_assets":
                return _dump(idx.search_assets(schema.IsccQuery(**args["query"]), limit=args["limit"], exact=args.get("exact", False)))
            if op == "len":
                return len(idx)
            raise AssertionError(op)

        if op == "reopen":
            idx.close()
            idx = B200Index(tmp_path / "flow", max_dim=256, **index_kwargs)
            assert {"assets": len(idx), "realm_id": idx._realm_id} == step["result"], where
            continue
        if "error" in step:
            with pytest.raises(ERRORS[step["error"]]) as ei:
                call()
            assert str(ei.value) == step["message"], where
            continue
        got, exp = call(), step["result"]
        if op == "search_assets":
            assert got["query"] == exp["query"], where
            _same_global(got["global_matches"], exp["global_matches"], args["limit"])
            assert got["chunk_matches"] == exp["chunk_matches"], where
        else:
            assert got == exp, where
    idx.close()
    return counts
