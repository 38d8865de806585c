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

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """/root/reference in the build container; on the benchmark box the git-ignored copy __graft_entry__.build() left in
    baseline/_ref (it travels with the tree like the built .so files do)."""
    for cand in (os.environ.get("PRIFIT_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "src")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()

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


def load(force_cpu=False):
    """Import the reference hot-path modules; returns a namespace of them.  force_cpu: make the reference's hard-coded
    `.cuda()` calls the identity even on a machine that has a GPU (the CPU arm of the benchmark)."""
    import torch

    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for name in _STUB_NAMES:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if force_cpu or not torch.cuda.is_available():
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
