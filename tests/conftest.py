"""pytest configuration: marker registration, import paths, shared fixtures."""
import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_PARENT = os.path.join(ROOT, "scikit-dsp-comm_b200")
for p in (ROOT, PKG_PARENT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def filters():
    with np.load(os.path.join(GOLDEN, "filters.npz")) as f:
        return {k: f[k] for k in f.files}


def _load_cases():
    out = []
    with np.load(os.path.join(GOLDEN, "ref_cases.npz")) as f:
        keys = sorted(k[:-2] for k in f.files if k.endswith("_x"))
        for k in keys:
            meta = ast.literal_eval(str(f[k + "_meta"]))
            out.append(dict(name=k, x=f[k + "_x"], y=f[k + "_y"], **meta))
    return out


_CASES = None


def ref_cases():
    global _CASES
    if _CASES is None:
        _CASES = _load_cases()
    return _CASES


@pytest.fixture(scope="session")
def cases():
    return ref_cases()


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The C oracle is the checker; build it once (gcc, <1 s)."""
    import oracle
    try:
        oracle.build()
    except Exception as e:                       # pragma: no cover
        print("oracle C build failed, numpy/pure-Python oracle only:", e)


@pytest.fixture(autouse=True)
def _seed_everything():
    """Deterministic inputs: every test starts from the same torch / numpy global seeds."""
    np.random.seed(100)
    try:
        import torch
        torch.manual_seed(100)
    except Exception:
        pass
    yield
