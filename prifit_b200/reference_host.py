"""Hosting the reference's own Python tree on top of this package.

The reference is a directory of scripts (no setup.py): `models/`, `src/`, `convex_loss.py`, the training scripts.  A
maintainer switches its hot path to this package with

    import prifit_b200.reference_host as host
    host.activate("/path/to/prifit")          # sys.path + prifit_b200.install() + pointnet kernels bound
    from models.pointnet2_part_seg_msg import get_model      # the reference's model, unmodified

`activate` (1) puts the tree on sys.path, (2) registers inert stand-ins for the GUI / mesh packages the reference
imports at module top but never needs on the training path and that may be absent (open3d, trimesh, matplotlib, ipdb,
transforms3d, lap, tensorboard_logger, torch_scatter) -- only for names that do not import, (3) calls
prifit_b200.install() so `src.mean_shift`, `src.ellipsoid_utils`, `src.ellipsoid_fitting`, `src.fitting_utils`,
`src.guard` and `convex_loss` resolve to this package, and (4) binds the PointNet++ geometric kernels into the
reference's models.pointnet_util.

Where the tree is looked for: the argument, $PRIFIT_REFERENCE_ROOT, <repo>/baseline/_ref (the git-ignored copy
__graft_entry__.build() makes for the benchmark box), /root/reference.
"""
import importlib
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_OPTIONAL = ["open3d", "trimesh", "ipdb", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "transforms3d",
             "transforms3d.affines", "transforms3d.euler", "lap", "tensorboard_logger", "torch_scatter"]


class _Inert(types.ModuleType):
    """Stand-in for an absent optional package: attribute access and calls return more stand-ins."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
        self.__all__ = ["utility", "geometry", "visualization"]          # `from open3d import *` (src/utils.py:2,14)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        child = _Inert(self.__name__ + "." + item)
        setattr(self, item, child)
        return child

    def __call__(self, *a, **k):
        return _Inert(self.__name__ + "()")


def find_tree(root=None):
    for cand in (root, os.environ.get("PRIFIT_REFERENCE_ROOT"), os.path.join(_HERE, "..", "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "convex_loss.py")) and os.path.isdir(os.path.join(cand, "models")):
            return os.path.abspath(cand)
    return None


def stub_optional_packages():
    for name in _OPTIONAL:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Inert(name)


def activate(root=None, bind_pointnet=True):
    """Returns the tree's path.  Idempotent."""
    import prifit_b200

    tree = find_tree(root)
    if tree is None:
        raise FileNotFoundError("no PRIFIT reference tree found (argument, $PRIFIT_REFERENCE_ROOT, baseline/_ref, /root/reference)")
    stub_optional_packages()
    if tree not in sys.path:
        sys.path.insert(0, tree)
    importlib.invalidate_caches()
    prifit_b200.install()
    if bind_pointnet:
        prifit_b200.bind_pointnet_ops(importlib.import_module("models.pointnet_util"))
    return tree
