"""prifit_b200 -- B200-native (sm_100a) mean-shift clustering + weighted ellipsoid fitting, forward and
backward, behind the Python call surface of Hippogriff/prifit's src/mean_shift.py,
src/ellipsoid_utils.py, src/ellipsoid_fitting.py, src/fitting_utils.py, src/guard.py and
convex_loss.py.  See DESIGN.md / INTEGRATION.md.

Importing the package does not need a GPU; the first operator call loads libprifit_b200.so and
fails loudly if it is missing (there is no CPU fallback).
"""
import importlib
import sys

__version__ = "0.1.0"

_MIRRORS = {
    "src.guard": "prifit_b200.guard",
    "src.mean_shift": "prifit_b200.mean_shift",
    "src.fitting_utils": "prifit_b200.fitting_utils",
    "src.ellipsoid_fitting": "prifit_b200.ellipsoid_fitting",
    "src.ellipsoid_utils": "prifit_b200.ellipsoid_utils",
    "convex_loss": "prifit_b200.convex_loss",
}


def install():
    """Make the reference's import paths (`from src.mean_shift import MeanShift`, `import convex_loss`,
    ...) resolve to this package, so the reference's training scripts run on the new path unmodified.
    Call before the reference's hot-path modules are imported.  The reference's own `src` package stays
    importable: every module that is not mirrored here (src.utils, src.VisUtils, src.sample_ellipsoid,
    src.augment_utils, ...) still comes from the reference tree on sys.path."""
    import types

    if "src" not in sys.modules:
        try:
            importlib.import_module("src")                 # the reference's package (namespace or regular), if reachable
        except ImportError:
            pkg = types.ModuleType("src")                  # no reference tree on sys.path: mirrored modules only
            pkg.__path__ = []
            sys.modules["src"] = pkg
    for ref_name, ours in _MIRRORS.items():
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        if ref_name.startswith("src."):
            setattr(sys.modules["src"], ref_name.split(".", 1)[1], mod)


def bind_pointnet_ops(pointnet_util_module):
    """Replace the geometric operators of the reference's models/pointnet_util.py (farthest_point_sample,
    query_ball_point and the 3-NN interpolation inside PointNetFeaturePropagation) with the device kernels of
    csrc/pointnet.cu, in place, on the module object the reference's models import."""
    from . import pointnet_util as ours

    ours.bind(pointnet_util_module)
