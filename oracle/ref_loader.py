"""Import the UNMODIFIED reference (SMSD75/Timetuning) in place from /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing under ``timetuning_b200/`` may import this.
Used by ``oracle/make_golden.py`` (fixture generation, build container only), by the
tests that cross-check against the live reference, and by ``bench.py``'s reference arm.
``/root/reference`` does not exist on the GPU box: there the git-ignored copy
``baseline/_ref/`` made by ``oracle/make_ref.py`` (it travels with the ``gpurun``
snapshot) is used; callers must check :func:`available` first and skip cleanly.

Recipe: the reference's modules import eight third-party roots that
are absent here and never touched by the hot path; they are replaced by MagicMock
packages.  ``anyio.maybe_async`` (removed in anyio 4) is re-added as identity.
No reference source is copied or modified.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIPPED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")       # oracle/make_ref.py (git-ignored copy)


def _find_root() -> str:
    env = os.environ.get("TIMET_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _SHIPPED):
        if os.path.isfile(os.path.join(cand, "mask_propagation.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()
_STUB_ROOTS = ("timm", "faiss", "skimage", "mmcv", "matplotlib", "nbformat",
               "pytorch_lightning", "torchmetrics", "wandb", "tensorboard")
_loaded: dict[str, types.ModuleType] = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mask_propagation.py"))


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serve MagicMock packages for the missing third-party roots (and submodules)."""

    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS and not _really_importable(root):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        return None


_real_cache: dict[str, bool] = {}


def _really_importable(root: str) -> bool:
    if root not in _real_cache:
        finder_backup = [f for f in sys.meta_path if not isinstance(f, _StubFinder)]
        spec = None
        for f in finder_backup:
            try:
                spec = f.find_spec(root, None)
            except Exception:
                spec = None
            if spec is not None:
                break
        # tensorboard/wandb exist in the build container but pull heavy deps; the hot
        # path never uses them, so a stub is always acceptable.
        _real_cache[root] = spec is not None and root not in ("wandb",)
    return _real_cache[root]


def load():
    """Return (my_utils, mask_propagation, time_tuning, dino_vision_transformer) reference modules."""
    if _loaded:
        return (_loaded["my_utils"], _loaded["mask_propagation"], _loaded["time_tuning"],
                _loaded["dino_vision_transformer"])
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    import anyio
    if not hasattr(anyio, "maybe_async"):
        anyio.maybe_async = lambda x: x
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch  # noqa: F401  (the reference seeds torch at import, time_tuning.py:66-69)
    for name in ("my_utils", "mask_propagation", "time_tuning", "dino_vision_transformer"):
        _loaded[name] = importlib.import_module(name)
    return load()


class _FE:
    """Duck-typed stand-in for models.FeatureExtractor (mask_propagation.py:402-405 reads
    ``.spatial_resolution``; time_tuning.py:85 reads ``.feature_dim``)."""

    def __init__(self, spatial_resolution, feature_dim=256):
        self.spatial_resolution = spatial_resolution
        self.feature_dim = feature_dim


def fake_feature_extractor(spatial_resolution, feature_dim=256):
    return _FE(spatial_resolution, feature_dim)
