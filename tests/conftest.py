"""pytest config: `gpu` marker = needs a real B200 (run by the driver with `-m gpu`)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (sm_100a)")
    config.addinivalue_line("markers", "refsrc: test needs /root/reference (build container only)")


@pytest.fixture(scope="session", autouse=True)
def _torch_threads():
    import torch

    torch.set_num_threads(min(4, os.cpu_count() or 1))
    yield


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests skip (instead of failing inside the first kernel launch) on a box without a CUDA device."""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
