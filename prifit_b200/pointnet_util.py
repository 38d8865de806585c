"""Device kernels for the geometric operators of the reference's models/pointnet_util.py (the stage in front of the
fitting path) and the hook that binds them into the reference's module.

    farthest_point_sample           reference :63-84     csrc/pointnet.cu fps_kernel (one CTA per cloud)
    query_ball_point                reference :87-107    ball_query_kernel (no N-long sort per query)
    three_interpolate               reference :287-294   three_nn_kernel + interpolate kernels (no S-long sort per point)
    bind(module)                    patches those three into the reference's module object (its grouping / gather helpers
                                    and its nn.Modules stay the reference's own code)

Index semantics are the reference's (first index on ties; ball query pads with the first hit).  CUDA fp32 tensors only.
"""
import torch

from . import _lib, ops
from .ops import _ptr, _stream


def farthest_point_sample(xyz, npoint, start=None):
    """xyz[B,N,3] -> int64 [B,npoint].  `start` (not in the reference's signature) fixes the first centroid; by default it
    is drawn like the reference does, torch.randint(0, N, (B,)) on the host generator (:74)."""
    xyz = ops._chk(xyz)
    B, N, _ = xyz.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.long)
    start = start.to(device=xyz.device, dtype=torch.long).contiguous()
    out = torch.empty(B, npoint, dtype=torch.long, device=xyz.device)
    _lib.call("prifit_fps", _ptr(xyz), _ptr(start), B, N, int(npoint), _ptr(out), _stream())
    return out


def query_ball_point(radius, nsample, xyz, new_xyz):
    """xyz[B,N,3], new_xyz[B,S,3] -> int64 [B,S,nsample]."""
    xyz, new_xyz = ops._chk(xyz), ops._chk(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty(B, S, nsample, dtype=torch.long, device=xyz.device)
    _lib.call("prifit_ball_query", _ptr(xyz), _ptr(new_xyz), B, N, S, float(radius), int(nsample), _ptr(out), _stream())
    return out


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points2, idx, weight):
        points2 = ops._chk(points2)
        B, S, D = points2.shape
        N = idx.shape[1]
        out = torch.empty(B, N, D, dtype=torch.float32, device=points2.device)
        _lib.call("prifit_interpolate_fwd", _ptr(points2), _ptr(idx), _ptr(weight), B, N, S, D, _ptr(out), _stream())
        ctx.save_for_backward(idx, weight)
        ctx.shape = (B, N, S, D)
        return out

    @staticmethod
    def backward(ctx, gout):
        idx, weight = ctx.saved_tensors
        B, N, S, D = ctx.shape
        g = torch.zeros(B, S, D, dtype=torch.float32, device=gout.device)
        gout = gout.contiguous()                         # held in a local: the library gets raw pointers
        _lib.call("prifit_interpolate_bwd", _ptr(gout), _ptr(idx), _ptr(weight), B, N, S, D, _ptr(g), _stream())
        return g, None, None


def three_nn(xyz1, xyz2):
    """xyz1[B,N,3], xyz2[B,S,3] -> (idx int32 [B,N,3], weight [B,N,3]) of reference :287-293."""
    xyz1, xyz2 = ops._chk(xyz1), ops._chk(xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    idx = torch.empty(B, N, 3, dtype=torch.int32, device=xyz1.device)
    weight = torch.empty(B, N, 3, dtype=torch.float32, device=xyz1.device)
    _lib.call("prifit_three_nn", _ptr(xyz1), _ptr(xyz2), B, N, S, _ptr(idx), _ptr(weight), _stream())
    return idx, weight


def three_interpolate(xyz1, xyz2, points2):
    """The interpolation of PointNetFeaturePropagation.forward (:283-295): xyz1[B,N,3], xyz2[B,S,3], points2[B,S,D] ->
    [B,N,D]; S == 1 repeats the single feature row.  Differentiable w.r.t. points2."""
    if xyz2.shape[1] == 1:
        return points2.repeat(1, xyz1.shape[1], 1)
    idx, weight = three_nn(xyz1, xyz2)
    return _Interpolate.apply(points2, idx, weight)


def _feature_propagation_forward(self, xyz1, xyz2, points1, points2):
    """Drop-in for PointNetFeaturePropagation.forward (reference :275-302): channel-first in and out, the 3-NN
    inverse-distance interpolation on the device kernels, the module's own 1x1 conv / batch-norm stack unchanged."""
    feats = three_interpolate(xyz1.transpose(1, 2), xyz2.transpose(1, 2), points2.transpose(1, 2).contiguous())
    x = feats.transpose(1, 2)
    if points1 is not None:
        x = torch.cat([points1, x], dim=1)
    for conv, bn in zip(self.mlp_convs, self.mlp_bns):
        x = torch.relu(bn(conv(x)))
    return x


def bind(module):
    """Patch the reference's models.pointnet_util module object in place: its farthest_point_sample / query_ball_point
    and the interpolation inside PointNetFeaturePropagation run on the kernels above.  Returns the module."""
    module.farthest_point_sample = farthest_point_sample
    module.query_ball_point = query_ball_point
    module.PointNetFeaturePropagation.forward = _feature_propagation_forward
    module._prifit_b200_bound = True
    return module
