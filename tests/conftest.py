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


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Strict-bound accounting of the parity helpers (tests/parity.py): how many compared elements fell
    outside 1e-4 abs (images) / 1e-3 of the tensor's scale (gradients), per kind."""
    try:
        import parity
    except Exception:
        return
    log = parity.STRICT_LOG
    if not log:
        return
    import json

    for kind in ("image", "grad"):
        rows = [r for r in log if r["kind"] == kind]
        if not rows:
            continue
        n = sum(r["n"] for r in rows)
        out = sum(r["outside_strict"] for r in rows)
        dirty = [r for r in rows if r["outside_strict"]]
        worst = max((r["worst"] for r in rows), default=0.0)
        terminalreporter.write_line(
            f"parity[{kind}]: {len(rows)} comparisons, {n} elements, {out} outside the strict bound "
            f"({out / max(n, 1):.2e}) in {len(dirty)} comparisons; worst {'abs' if kind == 'image' else 'rel-to-scale'} "
            f"error {worst:.3e}")
    path = os.environ.get("B200SPLAT_PARITY_LOG")
    if path:
        with open(path, "w") as f:
            for r in log:
                f.write(json.dumps(r) + "\n")
