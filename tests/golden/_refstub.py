"""Import harness that lets the UNMODIFIED reference (``/root/reference``) run in the build
container, where matplotlib / scikit-image / hydra / pytorch_lightning ... are absent.

Test infrastructure only.  It is used by ``tests/golden/make_golden.py`` (run by hand in the build
container, outputs committed under ``tests/golden/``) and by the optional
``tests/test_oracle_vs_reference.py`` which is skipped when ``/root/reference`` is absent (GPU box).
Nothing in the product package imports this file.

What it does:
* puts ``/root/reference`` and ``/root/reference/vital`` on ``sys.path``;
* a meta-path finder answers every import of an absent third-party package with a ``MagicMock``
  module, EXCEPT ``skimage.draw.line`` (a faithful Bresenham, restated from
  ``skimage/draw/_draw.pyx::_line``) and ``strenum`` (a real str-valued Enum), because results
  depend on them;
* RNG capture helpers so the reference's random draws can be recorded and replayed.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types
from contextlib import contextmanager
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = os.environ.get("CU_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = (
    "matplotlib", "mpl_toolkits", "skimage", "h5py", "hydra", "omegaconf", "pytorch_lightning",
    "comet_ml", "seaborn", "medpy", "uncertainties", "torchinfo", "pathos", "SimpleITK",
    "keras_preprocessing", "scripts", "dotenv", "torchmetrics", "torchvision", "albumentations",
    "nibabel", "tqdm_joblib", "PIL_stub_never",
)
_STUB_EXACT_PREFIXES = ("contour_uncertainty.data.ultromics",)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "contour_uncertainty"))


def bresenham_line(r0, c0, r1, c1):
    """``skimage.draw.line`` (skimage/draw/_draw.pyx::_line), restated.  Returns (rr, cc)."""
    r0, c0, r1, c1 = int(r0), int(c0), int(r1), int(c1)
    steep = 0
    r, c = r0, c0
    dr, dc = abs(r1 - r0), abs(c1 - c0)
    sc = 1 if (c1 - c) > 0 else -1
    sr = 1 if (r1 - r) > 0 else -1
    if dr > dc:
        steep = 1
        c, r = r, c
        dc, dr = dr, dc
        sc, sr = sr, sc
    d = 2 * dr - dc
    rr = np.zeros(max(dc, dr) + 1, dtype=np.intp)
    cc = np.zeros(max(dc, dr) + 1, dtype=np.intp)
    for i in range(dc):
        if steep:
            rr[i], cc[i] = c, r
        else:
            rr[i], cc[i] = r, c
        while d >= 0:
            r = r + sr
            d = d - 2 * dc
        c = c + sc
        d = d + 2 * dr
    rr[dc], cc[dc] = r1, c1
    return rr, cc


class _MockModule(types.ModuleType):
    """Module whose every missing attribute is a MagicMock (so ``from x import y`` works)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        mod = _MockModule(spec.name)
        mod.__path__ = []  # behaves as a package so sub-imports reach the finder again
        return mod

    def exec_module(self, module):
        name = module.__name__
        if name == "skimage.draw":
            module.line = bresenham_line
            # restated point-in-polygon fill (scikit-image is absent): parity unpinned, see oracle/port.py
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
            from oracle.port import draw_polygon

            module.polygon = draw_polygon
        if name == "skimage.measure":
            # restated marching squares (scikit-image is absent): parity unpinned, see oracle/port.py
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
            from oracle.port import find_contours

            module.find_contours = find_contours
        if name == "strenum":
            import enum

            class LowercaseStrEnum(str, enum.Enum):
                def _generate_next_value_(name, start, count, last_values):  # noqa: N805
                    return name.lower()

                def __str__(self):
                    return str(self.value)

            class SnakeCaseStrEnum(LowercaseStrEnum):
                pass

            module.LowercaseStrEnum = LowercaseStrEnum
            module.SnakeCaseStrEnum = SnakeCaseStrEnum
            module.StrEnum = LowercaseStrEnum
        if name == "pytorch_lightning":
            import torch.nn as nn

            class LightningModule(nn.Module):
                def save_hyperparameters(self, *a, **k):
                    pass

            module.LightningModule = LightningModule
            module.Callback = object
            module.LightningDataModule = object


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS or root == "strenum" or any(
            fullname == p or fullname.startswith(p + ".") for p in _STUB_EXACT_PREFIXES
        ):
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        return None


_installed = False


def install():
    """Idempotently install the stub finder and the reference paths."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (os.path.join(REFERENCE_ROOT, "vital"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.meta_path.append(_StubFinder())  # after the real finders: only absent packages are stubbed
    import torch

    torch.set_num_threads(1)
    _installed = True


@contextmanager
def replay_standard_normal(draws):
    """Feed ``torch.distributions.MultivariateNormal.rsample`` from ``draws`` (iterable of arrays).

    The reference's PSM sampler draws through ``MultivariateNormal(...).rsample((n,))`` which calls
    ``torch.distributions.multivariate_normal._standard_normal(shape, dtype, device)``.
    """
    import torch
    from torch.distributions import multivariate_normal as mvn

    it = iter(draws)
    orig = mvn._standard_normal
    log = []

    def fake(shape, dtype, device):
        z = torch.as_tensor(np.asarray(next(it)), dtype=dtype, device=device).reshape(shape)
        log.append(tuple(shape))
        return z

    mvn._standard_normal = fake
    try:
        yield log
    finally:
        mvn._standard_normal = orig
