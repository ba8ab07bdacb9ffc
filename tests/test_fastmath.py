"""csrc/fastmath64.cuh -- the straight-line double exp / log of the BP_DEC throughput kernel -- against libm on the CPU
(the header is host-callable): relative error within 2.3e-16 on the decoders' argument ranges, IEEE special values."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fx_exp_and_fx_log_against_libm(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "fastmath_check"
    subprocess.run([gxx, "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "fastmath_check.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
