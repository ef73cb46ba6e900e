import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O   # test infrastructure: the CPU restatement of the reference shaders
    O.build()
    return O


@pytest.fixture(scope="session")
def vk():
    import vk_renderer_b200 as V
    return V


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
