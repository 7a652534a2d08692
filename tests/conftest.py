import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    out = {}
    for k in d.files:
        v = d[k]
        if v.dtype.kind in "US":
            out[k] = str(v)
        elif v.ndim == 0:
            out[k] = v.item()
        else:
            out[k] = torch.from_numpy(v)
    return out


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Make sure the in-tree libraries exist (no-op when they are already built)."""
    from splat_one_b200 import build as b

    if not b.LIB.exists():
        b.build()
    from oracle import raster_ref

    raster_ref.build()
