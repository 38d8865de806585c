"""Intersection penalty between the fitted ellipsoids (reference convex_loss.py:346-441), on the device kernels of
csrc/intersect.cu.

    intersection_loss(params, points, version=3)   version 3 = compute_intersection_loss_volume_3 (what convex_loss calls at
                                                   :97, as its code is written with torch_scatter.scatter_mean), version 4 =
                                                   compute_intersection_loss_volume_4
    probe_points(chamfer_points)                   chamfer cloud minus a U[0, 0.2) jitter drawn like the reference draws it
                                                   (torch.rand on the host generator, then moved to the device, :97)

Note: the reference's shipped convex_loss.py raises NameError inside version 3 (the torch_scatter import is commented out
at :17); this module computes what that code computes once the import is restored.
"""
import os

import torch

from . import _lib, ops
from .ops import _ptr, _stream


class IntersectLoss(torch.autograd.Function):
    """(Q[B,M,3], s, V, c, valid, K, version) -> (loss_b[B], counted[B]); gradients flow to s, V, c."""

    @staticmethod
    def forward(ctx, Q, s, V, c, valid, K, version):
        Q, s, V, c = ops._chk(Q), ops._chk(s), ops._chk(V), ops._chk(c)
        B, M, _ = Q.shape
        kcap = s.shape[1]
        dev = Q.device
        nbytes = max(16, _lib.load().prifit_intersect_workspace_bytes(B, M))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        loss_b = torch.empty(B, dtype=torch.float32, device=dev)
        counted = torch.empty(B, dtype=torch.float32, device=dev)
        kstar = torch.empty(B, M, dtype=torch.int32, device=dev)
        aux = torch.empty(B, M, dtype=torch.float32, device=dev)
        _lib.call("prifit_intersect_fwd", _ptr(Q), _ptr(s), _ptr(V), _ptr(c), _ptr(valid), _ptr(K), B, M, kcap, int(version),
                  _ptr(loss_b), _ptr(counted), _ptr(kstar), _ptr(aux), _ptr(ws), nbytes, _stream())
        ctx.save_for_backward(Q, s, V, c, valid, K, kstar, aux)
        ctx.version = int(version)
        ctx.mark_non_differentiable(counted)
        return loss_b, counted

    @staticmethod
    def backward(ctx, gloss, _gcounted):
        Q, s, V, c, valid, K, kstar, aux = ctx.saved_tensors
        B, M, _ = Q.shape
        gs, gV, gc = torch.empty_like(s), torch.empty_like(V), torch.empty_like(c)
        gloss = gloss.contiguous()                       # held in a local: the library gets raw pointers
        _lib.call("prifit_intersect_bwd", _ptr(Q), _ptr(s), _ptr(V), _ptr(c), _ptr(valid), _ptr(K), _ptr(kstar), _ptr(aux),
                  _ptr(gloss), B, M, s.shape[1], ctx.version, _ptr(gs), _ptr(gV), _ptr(gc), _stream())
        return None, gs, gV, gc, None, None, None


def default_version():
    return int(os.environ.get("PRIFIT_INTERSECT_VERSION", "3"))


def probe_points(chamfer_points):
    """chamfer_points[B,M,3] -> chamfer_points - U[0, 0.2): `torch.rand(shape).cuda() * 0.2` of reference :97, i.e. one draw
    of B*M*3 numbers from the HOST generator (kept for generator-state parity), uploaded."""
    jitter = torch.rand(chamfer_points.shape).to(chamfer_points.device, non_blocking=True)
    return chamfer_points - jitter * 0.2


def intersection_loss(params, points, version=None):
    """params: ellipsoid_fitting.ParamsBatch (or a list of per-shape lists of (s, V, c)); points[B,M,3] probe points.
    Returns the scalar penalty (mean over the shapes with at least two ellipsoids, 0 if there is none)."""
    from .utils import _pad_params

    version = default_version() if version is None else int(version)
    s, V, c, valid, K = _pad_params(params, points.device)
    loss_b, counted = IntersectLoss.apply(points.contiguous(), s, V, c, valid, K, version)
    return (loss_b * counted).sum() / counted.sum().clamp(min=1.0)
