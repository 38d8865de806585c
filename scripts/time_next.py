#!/usr/bin/env python
"""Kernel times of the 'next' rows: entropy regulariser (N/4 sub-sample of 24 x 2048 x 128) and the
nearest-neighbour half of the analytic chamfer distance (24 shapes, 10000 sampled points vs 5000-point clouds)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    X = torch.nn.functional.normalize(torch.randn(24, 2048, 128, device=dev), dim=2).requires_grad_(True)
    idx = torch.from_numpy(np.random.RandomState(0).choice(2048, 512, replace=False).astype(np.int32)).to(dev)
    print("entropy fwd       %7.1f us" % timeit(lambda: ops.EntropyLoss.apply(X, idx)))
    print("entropy fwd+bwd   %7.1f us" % timeit(lambda: ops.EntropyLoss.apply(X, idx).sum().backward()))
    S = (torch.rand(24, 10000, 3, device=dev) * 2 - 1).requires_grad_(True)
    T = torch.rand(24, 5000, 3, device=dev) * 2 - 1
    print("nearest fwd       %7.1f us  (24 x 10000 x 5000 = %.2f G pairs)" % (timeit(lambda: ops.NearestSqDist.apply(S, None, T)), 24 * 1e4 * 5e3 / 1e9))
    print("nearest fwd+bwd   %7.1f us" % timeit(lambda: ops.NearestSqDist.apply(S, None, T)[0].sum().backward()))


def pointnet():
    from prifit_b200 import pointnet_util as pu
    dev = torch.device("cuda:0")
    xyz = torch.rand(24, 2048, 3, device=dev) * 2 - 1
    start = torch.zeros(24, dtype=torch.long)
    print("fps 2048 -> 512   %7.1f us" % timeit(lambda: pu.farthest_point_sample(xyz, 512, start=start)))
    new_xyz = torch.gather(xyz, 1, pu.farthest_point_sample(xyz, 512, start=start).unsqueeze(-1).expand(-1, -1, 3))
    print("ball query r=0.2  %7.1f us" % timeit(lambda: pu.query_ball_point(0.2, 32, xyz, new_xyz)))
    feats = torch.randn(24, 512, 128, device=dev, requires_grad=True)
    print("3-NN interp fwd   %7.1f us" % timeit(lambda: pu.three_interpolate(xyz, new_xyz, feats)))
    print("3-NN interp f+b   %7.1f us" % timeit(lambda: pu.three_interpolate(xyz, new_xyz, feats).sum().backward()))


def full_chamfer():
    import prifit_b200.convex_loss as cl
    from prifit_b200 import synthetic
    dev = torch.device("cuda:0")
    E, P, _ = synthetic.planted_shapes(24, n_points=2048, n_clusters=16, seed=1000)
    X = E.permute(0, 2, 1).contiguous().to(dev)
    Pc = P.permute(0, 2, 1).contiguous().to(dev)

    def step(full):
        Xi = X.detach().requires_grad_(True)
        total, _, _, _ = cl.convex_loss(Pc, Pc, Xi, quantile=0.05, iterations=10, max_num_clusters=25, full_chamfer=full)
        total.backward()

    print("convex_loss + backward, SDF half (graph path)       %7.1f us" % timeit(lambda: step(False), n=10))
    print("convex_loss + backward, full chamfer (eager path)   %7.1f us" % timeit(lambda: step(True), n=10))


if __name__ == "__main__":
    main()
    pointnet()
    full_chamfer()
