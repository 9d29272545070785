"""Opt-in drop-in alias: after ``install_alias()`` the reference's import paths resolve to this package,

    import contouring_uncertainty_b200.compat as compat; compat.install_alias()
    from contour_uncertainty.task.regression.dsnt.dsnt_al import DSNTAleatoric        # -> contouring_uncertainty_b200...
    from contour_uncertainty.sampler.posterior_shape_model.psm import PosteriorShapeModelSampler

so a caller of the predict path changes two lines, not its imports.  Only modules this package mirrors exist (the predict
path of SURVEY.md section 8); importing anything else of the reference (training losses, datasets, plotting ...) raises
ModuleNotFoundError, loudly.  Do not install the alias in a process that also imports the real reference.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys

ALIAS = "contour_uncertainty"
TARGET = __name__.rsplit(".", 1)[0]        # contouring_uncertainty_b200


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        mod = importlib.import_module(self.target)       # the very same module object under a second name
        self.real_spec = mod.__spec__
        return mod

    def exec_module(self, module):
        module.__spec__ = self.real_spec                 # the import machinery stamped the alias spec on it: put the real one back


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != ALIAS and not fullname.startswith(ALIAS + "."):
            return None
        real = TARGET + fullname[len(ALIAS):]
        try:
            mod = importlib.import_module(real)
        except ModuleNotFoundError as e:
            if e.name is not None and not e.name.startswith(TARGET):
                raise          # a module this package mirrors exists but one of ITS imports is missing: say which
            return None                                   # not mirrored: the import fails as ModuleNotFoundError
        spec = importlib.machinery.ModuleSpec(fullname, _AliasLoader(real), is_package=hasattr(mod, "__path__"))
        return spec


_finder = None


def install_alias() -> None:
    """Idempotent.  Refuses to shadow an already imported ``contour_uncertainty`` that is not this package."""
    global _finder
    have = sys.modules.get(ALIAS)
    if have is not None and getattr(have, "__name__", "") != TARGET:
        raise ImportError(f"{ALIAS} is already imported from {getattr(have, '__file__', '?')}: not shadowing it")
    if _finder is None:
        _finder = _AliasFinder()
        sys.meta_path.insert(0, _finder)


def uninstall_alias() -> None:
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for k in [k for k in sys.modules if k == ALIAS or k.startswith(ALIAS + ".")]:
        del sys.modules[k]
