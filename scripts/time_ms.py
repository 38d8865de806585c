#!/usr/bin/env python
"""Time the all-seed mean-shift kernel (tcgen05) on cfg2 / cfg4 shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import ops, synthetic  # noqa: E402


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
    for B, N in ((24, 2048), (16, 10000), (1, 2048)):
        E, _, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=0)
        X = ops.normalize_fwd(E.to(dev))
        bw = torch.full((B,), 0.15, device=dev)
        flops = 4.0 * N * N * 128 * 10 * B
        t = timeit(lambda: ops.meanshift(X, bw, 10, ops.MS_F16_TCGEN05))
        print("B=%d N=%d  %8.1f us  %.0f TFLOP/s" % (B, N, t, flops / t / 1e6), flush=True)


if __name__ == "__main__":
    main()
