#!/usr/bin/env python
"""Time membership forward / backward on a cfg2-shaped batch for different cluster sizes (PRIFIT_MEMB_CLUSTER)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import ops, pipeline, synthetic  # noqa: E402


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
    N = 2048
    for B in (8, 24):
        E, _, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=0)
        X = ops.normalize_fwd(E.to(dev))
        res = pipeline.cluster_batch(X, N, 0.05, 10, 25)
        _, _, C = ops.rows_fwd(X, res.bw, res.idx, res.K, 10, res.kcap)
        W, smax = ops.membership_fwd(C, X, res.bw, res.K)
        gW = torch.randn_like(W)
        gX = torch.zeros_like(X)
        tf = timeit(lambda: ops.membership_fwd(C, X, res.bw, res.K))
        print("B=%d  fwd %.1f us" % (B, tf), flush=True)
        tb = timeit(lambda: ops.membership_bwd(C, X, res.bw, res.K, W, smax, gW, gX))
        print("B=%d  bwd %.1f us" % (B, tb), flush=True)


if __name__ == "__main__":
    main()
