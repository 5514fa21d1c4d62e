import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # -m gpu tests are only ever collected on a machine with a GPU; skip them loudly otherwise.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return load


# north-star tolerance for soft labels and Sinkhorn Q (BASELINE.json): |a-b| <= 1e-4 + 1e-5*|b|
ATOL, RTOL = 1e-4, 1e-5


def assert_close(actual, expected, atol=ATOL, rtol=RTOL, what=""):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, f"{what}: shape {actual.shape} vs {expected.shape}"
    err = np.abs(actual - expected)
    bad = err > atol + rtol * np.abs(expected)
    assert not bad.any(), (f"{what}: {bad.sum()} / {bad.size} outside tol, max abs err {err.max():.3e} "
                           f"at {np.unravel_index(err.argmax(), err.shape)}")


@pytest.fixture
def timet_env(monkeypatch):
    """Set TIMET_* experiment switches for one test: the library reads the environment once, so it is told to
    re-read after every change and once more when the test is over."""
    from timetuning_b200 import _cabi

    def set_env(**kv):
        for k, v in kv.items():
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, str(v))
        _cabi.reload_env()
    yield set_env
    monkeypatch.undo()
    _cabi.reload_env()
