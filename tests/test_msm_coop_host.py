"""Cooperative (four-role) XYZZ addition / doubling of csrc/msm_coop.cuh on the host: the level functions the device
kernels call are plain C++, so the four roles are run in lockstep by a small g++ harness (tests/host/coop_add_test.cpp)
and compared with curve.cuh::xyzz_add / xyzz_double, special cases included."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cooperative_addition_matches_the_serial_formulas():
    src = os.path.join(ROOT, "tests", "host", "coop_add_test.cpp")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "coop_add_test")
        subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "0g-halo2_b200", "csrc"), src, "-o", exe],
                       check=True, capture_output=True, text=True)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout
