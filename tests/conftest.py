import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "vit-lens_b200")
for p in (SRC, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def emu(monkeypatch):
    """Swap the engine's kernel gateway for the torch-CPU emulation (host-logic tests only)."""
    from vitlens_b200 import engine
    from tests import emu_ops

    monkeypatch.setattr(engine, "_ops", emu_ops)
    engine.WEIGHTS.clear()
    yield emu_ops
    engine.WEIGHTS.clear()
