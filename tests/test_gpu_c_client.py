"""Runs the plain-C client of include/gfi.h (tests/c_client/abi_client.c) on the GPU: the calls a Rust `impl Index`
makes, checked against the reference's known answers, with no Python between the client and libgfi."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_client_runs_the_trait_calls_against_libgfi():
    from test_abi_and_host import build_c_client
    exe = build_c_client()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_client ok" in out.stdout
