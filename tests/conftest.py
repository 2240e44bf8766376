import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "0g-halo2_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "models"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One zg_b200 context on cuda:0 bound to torch's current stream (gpu tests only)."""
    import torch
    import zg_b200
    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    c = zg_b200.Context(0, stream.cuda_stream)
    yield c
    c.close()
