"""B200-native (sm_100a) sparse-embedding + feature-interaction hot path of
PatrickHwang/Explicit-tf2-Recommendation, behind the reference's own layer API.

The directory name carries a hyphen (it mirrors the reference repository name),
so import it as ``import etr_b200`` (a tiny alias package at the repo root) or
``importlib.import_module("explicit-tf2-recommendation_b200")``.

Importing this package never touches CUDA; constructing a layer does, and fails
loudly if ``libetr.so`` is not built or no sm_100a device is present.
"""
from . import _lib  # noqa: F401
from ._lib import EtrError, EtrIdRangeError  # noqa: F401


_SUBMODULES = ("CustomLayers", "runtime", "dense", "sharded", "build", "autograd", "tf_adapter", "tf_checkpoint", "tfrecord")


def __getattr__(name):
    # layers import torch + the runtime lazily so that `import etr_b200` stays cheap
    import importlib
    if name in _SUBMODULES:
        return importlib.import_module("." + name, __name__)
    if name.startswith("__"):
        raise AttributeError(name)
    mod = importlib.import_module(".CustomLayers", __name__)
    try:
        return getattr(mod, name)
    except AttributeError:
        rt = importlib.import_module(".runtime", __name__)
        if hasattr(rt, name):
            return getattr(rt, name)
        raise AttributeError(name)
