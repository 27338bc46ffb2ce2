import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ozlib():
    """The product library through its C-ABI (never the oracle)."""
    import ozimmu_b200
    return ozimmu_b200.lib()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.oracle()


@pytest.fixture(scope="session")
def ozref():
    """The unmodified reference built by oracle/Makefile into oracle/_ref/libozref.so (GPU only)."""
    import oracle_lib
    lib = oracle_lib.reference()
    if lib is None:
        pytest.skip("oracle/_ref/libozref.so not built")
    return lib


@pytest.fixture()
def handle():
    import ozimmu_b200 as oz
    h = oz.create()
    yield h
    oz.destroy(h)
