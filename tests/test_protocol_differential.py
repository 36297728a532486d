"""
Differential test against the LIVE reference: the reference's unmodified `UsearchIndex` is run here (loaded from
/root/reference by tests/golden/make_protocol_golden.py, with its three un-installable dependencies stubbed as that
script documents) on freshly seeded scenarios, and `B200Index` must reproduce every step. Skipped where the reference
tree does not exist (the GPU box); the committed fixture of one such run is replayed there instead.
"""

import importlib.util
import json
from pathlib import Path

import pytest

from tests.protocol_replay import replay

REF = Path("/root/reference")
GEN = Path(__file__).parent / "golden" / "make_protocol_golden.py"

pytestmark = pytest.mark.skipif(not (REF / "iscc_search").is_dir(), reason="the reference tree only exists in the build container")

_loaded = {}


def _generator():
    if not _loaded:
        spec = importlib.util.spec_from_file_location("make_protocol_golden", GEN)
        gen = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(gen)
        _loaded["gen"] = gen
        _loaded["ref"] = gen.load_reference()
    return _loaded["gen"], _loaded["ref"]


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_backend_reproduces_the_live_reference(seed, tmp_path, cpu_stores):
    gen, (index_mod, schema) = _generator()
    steps = json.loads(json.dumps(gen.run(index_mod, schema, tmp_path / "reference", seed=seed)))
    assert sum(1 for s in steps if s["op"] == "search_assets") > 100
    counts = replay(tmp_path / "ours", steps=steps)
    assert counts["add_assets"] >= 6
