"""Alias: ``import etr_b200`` == the package in ``explicit-tf2-recommendation_b200/``
(whose directory name is not a valid Python identifier).  ``etr_b200.X`` resolves
to the very same module object as ``explicit-tf2-recommendation_b200.X`` (a
meta-path finder maps the names), so there is exactly one copy of every class."""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys

_REAL = "explicit-tf2-recommendation_b200"
_ALIAS = __name__

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])     # the one real module object

    def exec_module(self, module):
        pass


if not any(type(f).__name__ == "_AliasFinder" for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
sys.modules[_ALIAS] = _pkg
