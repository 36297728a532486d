"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/isx.h declares, and fails loudly without a device."""

import ctypes
import re
from pathlib import Path

import pytest

from iscc_search_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_are_all_exported():
    header = (ROOT / "include" / "isx.h").read_text()
    declared = set(re.findall(r"\b(isx_[a-z0-9_]+)\s*\(", header))
    declared -= {"isx_store", "isx_stats"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.isx_abi_version() == 1


def test_no_torch_types_in_the_abi():
    header = (ROOT / "include" / "isx.h").read_text()
    code = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # signatures only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code


def test_library_is_built_for_sm_100a_only():
    mk = (ROOT / "iscc_search_b200" / "csrc" / "Makefile").read_text()
    assert "arch=compute_100a,code=sm_100a" in mk and "-lineinfo" in mk


def test_product_never_imports_the_oracle():
    for py in (ROOT / "iscc_search_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
        assert "import tests" not in src and "from tests" not in src, py  # the oracle-backed store double is test-only
    for src in (ROOT / "iscc_search_b200" / "csrc").iterdir():
        if src.suffix in (".cu", ".cuh", ".hpp"):
            assert "oracle" not in src.read_text().lower(), src


def _device_count():
    n = ctypes.c_int()
    rc = _lib.lib().isx_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


@pytest.mark.skipif(_device_count() > 0, reason="a GPU is present; this checks the no-device behaviour")
def test_open_without_device_fails_loudly_no_cpu_fallback(tmp_path):
    with pytest.raises(_lib.IsxError, match="no CPU fallback"):
        _lib.Store()
    from iscc_search_b200 import ShardedNphdIndex

    with pytest.raises(_lib.IsxError):
        ShardedNphdIndex(max_dim=256)
    from iscc_search_b200.backend import B200IndexManager
    from iscc_search_b200.schema import IsccIndex

    with pytest.raises(_lib.IsxError, match="no CPU fallback"):  # the protocol backend has no CPU path either
        B200IndexManager(tmp_path).create_index(IsccIndex(name="nogpu"))


def test_argument_validation_happens_before_any_device_work():
    h = ctypes.c_void_p()
    L = _lib.lib()
    assert L.isx_open(ctypes.byref(h), 0, 12, 32, 0) == _lib.ISX_EINVAL
    assert b"key_bytes" in L.isx_last_error()
    assert L.isx_open(ctypes.byref(h), 0, 8, 40, 0) == _lib.ISX_EINVAL
    assert L.isx_open(ctypes.byref(h), 0, 8, 16, 32) == _lib.ISX_EINVAL


def test_share_api_validates_arguments():
    L = _lib.lib()
    buf = (ctypes.c_ubyte * 64)()
    assert L.isx_share_init(None, 2, 0, 16, buf) == _lib.ISX_EINVAL
    assert L.isx_share_attach(None, 1, buf) == _lib.ISX_EINVAL
    assert L.isx_share_reset(None) == _lib.ISX_EINVAL


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """include/isx.h is C99 (what cgo / JNI / N-API bind), and examples/isx_client.c builds against the library with gcc alone.
    Without a CUDA device the client must stop at isx_open with the library's message - no CPU path behind the ABI."""
    import shutil
    import subprocess

    import torch

    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of this image"
    exe = tmp_path / "isx_client"
    so_dir = ROOT / "iscc_search_b200"
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "isx_client.c"),
           "-o", str(exe), f"-L{so_dir}", "-lisx_b200", f"-Wl,-rpath,{so_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        return  # with a device the client really searches: tests/test_gpu_c_client.py
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 2
    assert "isx_open failed" in run.stderr and "no CPU fallback" in run.stderr
