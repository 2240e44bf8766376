#!/usr/bin/env python
"""Regenerates tests/golden/kernels.json: SHA-256 digests of the CPU oracle's outputs for seeded inputs of every
per-function entry point of the C ABI.  `-m "not gpu"` re-derives them from the oracle
(drift guard for the checker), `-m gpu` checks the CUDA path against the same committed digests."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "0g-halo2_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import bn254  # noqa: E402
import cpu_ref  # noqa: E402
from golden_kernel_cases import CASES, digest  # noqa: E402


if __name__ == "__main__":
    out = {name: digest(case["oracle"]()) for name, case in CASES.items()}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "kernels.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))
