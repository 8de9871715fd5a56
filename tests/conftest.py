import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) where there is no CUDA device or the library has not been built, so the
    CPU-only suite's result is not masked; on the GPU box both exist and every gpu test runs."""
    import torch
    lib = os.path.join(ROOT, "monohair_b200", "libmonohair_b200.so")
    why = None
    if not torch.cuda.is_available():
        why = "no CUDA device"
    elif not os.path.exists(lib):
        why = "libmonohair_b200.so has not been built (python -m monohair_b200.build)"
    if why:
        skip = pytest.mark.skip(reason=why)
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)
