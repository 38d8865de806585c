#!/usr/bin/env python
"""Time the K-seed trajectory kernels (forward / backward, both engines) on a cfg2-shaped batch.
PRIFIT_ROWS_CLUSTER is read by the library at every launch, so cluster sizes are swept in-process."""
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
    B, N, kc = int(os.environ.get("B", 24)), int(os.environ.get("N", 2048)), int(os.environ.get("KC", 16))
    dev = torch.device("cuda:0")
    E, _, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=kc, seed=0)
    X = ops.normalize_fwd(E.to(dev))
    res = pipeline.cluster_batch(X, N, 0.05 if N < 5000 else 0.01, 10, 50 if N >= 5000 else 25)
    gC = torch.randn(B, res.kcap, 128, device=dev)
    gX = torch.zeros_like(X)
    print("B=%d N=%d K=%s kcap=%d" % (B, N, res.K_host[:4], res.kcap))
    for engine, name in ((1, "simt"), (0, "tc")):
        for cs in ([0] if engine == 1 else [int(c) for c in os.environ.get("CS_LIST", "8,4").split(",")]):
            if cs:
                os.environ["PRIFIT_ROWS_CLUSTER"] = str(cs)
            traj, stat, C = ops.rows_fwd(X, res.bw, res.idx, res.K, 10, res.kcap, engine)
            tf = timeit(lambda: ops.rows_fwd(X, res.bw, res.idx, res.K, 10, res.kcap, engine))
            tb = timeit(lambda: ops.rows_bwd(X, res.bw, res.idx, res.K, traj, stat, gC, gX, 10, res.kcap, engine))
            print("%-5s cluster=%d  fwd %8.1f us   bwd %8.1f us" % (name, cs, tf, tb), flush=True)


if __name__ == "__main__":
    main()
