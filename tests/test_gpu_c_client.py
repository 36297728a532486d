"""The C99 client of the C ABI (examples/isx_client.c) on a real device: no Python, no torch in that process."""

import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_c_client_searches_on_the_device(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc on this box")
    exe = tmp_path / "isx_client"
    so_dir = ROOT / "iscc_search_b200"
    build = subprocess.run([gcc, "-std=c99", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "isx_client.c"), "-o", str(exe),
                            f"-L{so_dir}", "-lisx_b200", f"-Wl,-rpath,{so_dir}"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stderr
    # 64-bit query against 8/16/8/32-byte codes: compared on the common 8-byte prefix, ties ordered by key
    assert run.stdout.split("\n")[:4] == ["key 12  nphd 0/64", "key 14  nphd 0/64", "key 13  nphd 2/64", "key 11  nphd 8/64"]
