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


# ------------------------------------------------------------------------------------------------------------------
# cfg5 (BASELINE.json configs[4]): the reference's PointNet++ MSG part-segmentation model, unmodified, on top of this
# package -- models/pointnet2_part_seg_msg.py:64-134 calls convex_loss at :110; the training loop follows
# train_partseg_shapenet.py:441-451 (self-supervised step: 2048 of the 5000 chamfer points in, loss = mean(total) * lmbda).
# ------------------------------------------------------------------------------------------------------------------
def build_partseg_model(device, num_parts=50, root=None, seed=0):
    """The reference's models.pointnet2_part_seg_msg.get_model(num_parts) (1.76 M parameters), instantiated from the
    hosted tree with this package's convex_loss / clustering / fitting behind it and the PointNet++ geometric kernels
    bound.  Returns (model, module)."""
    import torch

    activate(root)
    module = importlib.import_module("models.pointnet2_part_seg_msg")
    if module.convex_loss.__module__ != "prifit_b200.convex_loss":
        raise RuntimeError("the reference's model module was imported before prifit_b200.install(): it still calls its own convex_loss")
    torch.manual_seed(seed)
    model = module.get_model(num_parts).to(device)
    return model, module


def synthetic_partseg_batch(batch, n_points=2048, n_chamfer=5000, seed=0):
    """ShapeNet-shaped synthetic inputs on the HOST (the data loader's side of the step): chamfer_points[B,3,5000]
    (a union of anisotropic blobs scaled into the unit ball, like a normalised part-annotated shape), points[B,3,2048] =
    a random 2048-subset of them (train_partseg_shapenet.py:441), cls_label[B,16] one-hot category."""
    import numpy as np
    import torch

    from . import synthetic

    _, P, _ = synthetic.planted_shapes(batch, n_points=n_chamfer, n_clusters=8, seed=seed)
    P = P - P.mean(1, keepdim=True)
    P = P / P.norm(dim=2).amax(1).view(-1, 1, 1)
    chamfer = P.permute(0, 2, 1).contiguous()
    rs = np.random.RandomState(seed)
    choice = torch.from_numpy(rs.choice(n_chamfer, n_points, replace=False))
    points = chamfer[:, :, choice].contiguous()
    cls = torch.zeros(batch, 16)
    cls[torch.arange(batch), torch.from_numpy(rs.randint(0, 16, size=batch))] = 1.0
    return points, chamfer, cls


def partseg_selfsup_step(model, optimizer, points, chamfer, cls, quantile=0.05, msc_iterations=10, max_num_clusters=25, lmbda=1.0):
    """One self-supervised training step exactly as train_partseg_shapenet.py:444-451 drives the model: forward with
    include_convex_loss=True, ss_loss = mean(loss_self_sup) * lmbda, backward, optimizer step.  Returns the model's outputs
    (8-tuple, models/pointnet2_part_seg_msg.py:134) and the scalar loss."""
    import torch

    out = model(points, cls, chamfer_points=chamfer, include_convex_loss=True, quantile=quantile,
                msc_iterations=msc_iterations, max_num_clusters=max_num_clusters)
    ss_loss = torch.mean(out[3]) * lmbda
    optimizer.zero_grad(set_to_none=True)
    ss_loss.backward()
    optimizer.step()
    return out, ss_loss
