import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def maps():
    """name -> (IBSP, path); generated on first use, cached on disk."""
    from tools import mapgen
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = mapgen.get_map(name)
        return cache[name]
    return get


@pytest.fixture(scope="session")
def native():
    """the built product library (fails loudly when it is missing)."""
    import bsp_b200
    bsp_b200.build()
    bsp_b200.load_library()
    return bsp_b200


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)
