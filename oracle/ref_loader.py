"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference under /root/reference.

Only `oracle/make_golden.py` (run in the build container, where /root/reference is mounted)
uses this.  Nothing in the product package, the `-m gpu` tests, `smoke()` or `bench.py` may import
it: /root/reference does not exist on the GPU box.

The reference's hot-path modules import GUI / mesh packages that are absent here (open3d, trimesh,
matplotlib, ipdb, transforms3d, lap, tensorboard_logger) and hard-code `.cuda()`
(src/mean_shift.py:178,180, src/ellipsoid_fitting.py:38, src/fitting_utils.py:70-103).  We insert
inert stub modules and make `Tensor.cuda` the identity on a CPU-only host, then import the
reference modules as they are.
"""
import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PRIFIT_REFERENCE_ROOT", "/root/reference")

_STUB_NAMES = [
    "open3d", "trimesh", "ipdb", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
    "transforms3d", "transforms3d.affines", "transforms3d.euler", "lap", "tensorboard_logger",
    "torch_scatter",
]


class _Stub(types.ModuleType):
    """Module whose every attribute is another stub and which is callable."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
        self.__all__ = ["utility", "geometry", "visualization"]

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        child = _Stub(self.__name__ + "." + item)
        setattr(self, item, child)
        return child

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def load():
    """Import the reference hot-path modules; returns a namespace of them."""
    import torch

    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for name in _STUB_NAMES:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.guard = importlib.import_module("src.guard")
    ns.mean_shift = importlib.import_module("src.mean_shift")
    ns.fitting_utils = importlib.import_module("src.fitting_utils")
    ns.ellipsoid_fitting = importlib.import_module("src.ellipsoid_fitting")
    ns.ellipsoid_utils = importlib.import_module("src.ellipsoid_utils")
    ns.convex_loss = importlib.import_module("convex_loss")
    ns.utils = importlib.import_module("src.utils")
    return ns
