import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))
os.environ.setdefault("OMP_NUM_THREADS", "1")

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) where no CUDA device is visible."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _npz(name):
    return dict(np.load(GOLDEN / name, allow_pickle=False))


@pytest.fixture(scope="session")
def grooming_leg():
    return _npz("grooming_leg.npz")


@pytest.fixture(scope="session")
def grooming_head():
    return _npz("grooming_head.npz")


@pytest.fixture(scope="session")
def grooming_align():
    return _npz("grooming_align.npz")


@pytest.fixture(scope="session")
def locomotion():
    return _npz("locomotion.npz")


@pytest.fixture(scope="session")
def synthetic_gold():
    return _npz("synthetic.npz")


@pytest.fixture(scope="session")
def synthetic_long():
    return _npz("synthetic_long.npz")


@pytest.fixture(scope="session")
def generic_gold():
    return _npz("generic_leg.npz")


@pytest.fixture(scope="session")
def synthetic_wide():
    return _npz("synthetic_wide.npz")
