"""pytest configuration: the `gpu` marker, import paths, and loaders for the oracle / reference module."""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
for p in (str(ROOT), str(ROOT / "oracle"), str(ROOT / "oracle" / "_ref")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _try_import(name):
    try:
        return __import__(name)
    except Exception:  # noqa: BLE001 - absent or unloadable (no libcuda) is a legitimate state
        return None


@pytest.fixture(scope="session")
def ref_cpu():
    """The unmodified reference compiled in place (oracle/_ref/nerfpp_ref_cpu.so) or None when it was not built."""
    return _try_import("nerfpp_ref_cpu")


@pytest.fixture(scope="session")
def ref_cuda():
    import torch
    if not torch.cuda.is_available():
        return None
    return _try_import("nerfpp_ref_cuda")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(GOLDEN / name, allow_pickle=False)
    return load
