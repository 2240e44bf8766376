"""Committed digests tests/golden/kernels.json (scripts/make_golden_kernels.py) for every per-function entry point of the
C ABI.  CPU leg: the oracle still produces them.  GPU leg: libzg_b200.so produces the same bytes."""
import json
import os

import pytest

from golden_kernel_cases import CASES, digest

FIX = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kernels.json")))


def test_fixture_covers_every_case():
    assert sorted(FIX) == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_fixture(name):
    assert digest(CASES[name]["oracle"]()) == FIX[name]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_fixture(ctx, name):
    assert digest(CASES[name]["gpu"](ctx)) == FIX[name]
