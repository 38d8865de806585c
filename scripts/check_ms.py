#!/usr/bin/env python
"""All-seed mean-shift, tensor-core engine vs the fp32 CUDA-core engine (max |diff|), then timings.
Correctness (max |diff| on ragged sizes) and timings of the shipped kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import ops, synthetic  # noqa: E402
from scripts.time_ms import timeit  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for B, N in ((3, 2048), (2, 1100), (2, 128), (1, 100), (1, 10000)):
        E, _, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=8, seed=1)
        X = ops.normalize_fwd(E.to(dev))
        bw = torch.full((B,), 0.2, device=dev)
        a = ops.meanshift(X, bw, 10, ops.MS_F16_TCGEN05)
        torch.cuda.synchronize()
        r = ops.meanshift(X, bw, 10, ops.MS_FP32_SIMT)
        print("B=%d N=%d  max|tc - fp32| = %.3e  finite=%s" % (B, N, (a - r).abs().max().item(), bool(torch.isfinite(a).all())), flush=True)
    for B, N in ((24, 2048), (8, 2048), (16, 10000)):
        E, _, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=0)
        X = ops.normalize_fwd(E.to(dev))
        bw = torch.full((B,), 0.15, device=dev)
        flops = 4.0 * N * N * 128 * 10 * B
        t = timeit(lambda: ops.meanshift(X, bw, 10, ops.MS_F16_TCGEN05))
        print("B=%d N=%d  %8.1f us  %.0f TFLOP/s" % (B, N, t, flops / t / 1e6), flush=True)


if __name__ == "__main__":
    main()
