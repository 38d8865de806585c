"""Mirror of the geometric operators of the reference's models/pointnet_util.py (the stage in front of the fitting path).

    square_distance, index_points   reference :18-60     torch expressions (kept for API compatibility)
    farthest_point_sample           reference :63-84     csrc/pointnet.cu fps_kernel (one CTA per cloud)
    query_ball_point                reference :87-107    ball_query_kernel (no N-long sort per query)
    three_interpolate               reference :287-294   three_nn_kernel + interpolate kernels (no S-long sort per point)
    sample_and_group                reference :110-136

Index semantics are the reference's (first index on ties; ball query pads with the first hit).  CUDA fp32 tensors only.
"""
import torch

from . import _lib, ops
from .ops import _ptr, _stream


def square_distance(src, dst):
    """reference :18-41."""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def index_points(points, idx):
    """reference :44-60: points[B,N,C], idx[B,S...] -> [B,S...,C]."""
    B = points.shape[0]
    view_shape = [B] + [1] * (idx.dim() - 1)
    repeat_shape = [1] + list(idx.shape[1:])
    batch_indices = torch.arange(B, dtype=torch.long, device=points.device).view(view_shape).repeat(repeat_shape)
    return points[batch_indices, idx, :]


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


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False):
    """reference :110-136."""
    B, N, C = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, S, 1, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz_norm, index_points(points, idx)], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points
