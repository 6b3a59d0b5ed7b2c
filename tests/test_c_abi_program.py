"""A plain-C program linked against libmsfl.so exercises the C ABI end to end (no Python in the loop)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from msf_loam_b200 import _lib
    _lib.load_library()
    exe = str(tmp_path / "smoke")
    cmd = [shutil.which("gcc") or "gcc", "-std=c11", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "smoke.c"), "-L", os.path.join(ROOT, "msf_loam_b200"), "-lmsfl", "-lm",
           "-Wl,-rpath," + os.path.join(ROOT, "msf_loam_b200"), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_c_program_compiles_and_links_against_the_abi(tmp_path):
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_c_program_recovers_known_transform(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
