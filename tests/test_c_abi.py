# -*- coding: utf-8 -*-
"""The boundary is a C ABI: a plain C program (tests/c/abi_test.c) links libr48.so and the oracle
and compares the host-buffer entry points -- no Python or torch in the calling process."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_c_test(tmp_path):
    from rein48_b200 import _native
    from oracle import oracle
    _native.build()
    oracle.build()
    exe = str(tmp_path / "abi_test")
    pkg, orc = os.path.join(ROOT, "rein48_b200"), os.path.join(ROOT, "oracle")
    subprocess.check_call([
        "gcc", "-O1", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"),
        os.path.join(ROOT, "tests", "c", "abi_test.c"), "-o", exe,
        "-L", pkg, "-L", orc, "-l:libr48.so", "-l:libr48_oracle.so", "-lpthread",
        "-Wl,-rpath," + pkg, "-Wl,-rpath," + orc])
    return exe


def test_c_caller_compiles_and_links(tmp_path):
    """CPU box: the header is valid C and the library satisfies every symbol the C caller uses."""
    assert os.path.exists(build_c_test(tmp_path))


@pytest.mark.gpu
def test_c_caller_runs_bit_exact(tmp_path):
    out = subprocess.run([build_c_test(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "C ABI ok" in out.stdout
