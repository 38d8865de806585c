#!/usr/bin/env python
"""Kernel times of the bandwidth stage (gram_tc_kernel HIST x2 + COLLECT) and the NMS Gram passes at cfg2, from CUPTI.
PRIFIT_GRAM_DEBUG selects the timing experiments of gram_tc.cu (results are then meaningless)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import ops, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, N = int(os.environ.get("B", 24)), int(os.environ.get("N", 2048))
    E, _, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=0)
    X = ops.normalize_fwd(E.to(dev))
    kth = torch.full((B,), int(0.05 * N), dtype=torch.int32, device=dev)
    for _ in range(3):
        bw = ops.bandwidth(X, kth)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            bw = ops.bandwidth(X, kth)
            ops.nms(X, bw, 32)
        torch.cuda.synchronize()
    agg = {}
    for e in prof.profiler.kineto_results.events():
        if e.device_type() == torch.autograd.DeviceType.CUDA and "gram_tc" in e.name():
            agg.setdefault(e.name()[40:75], []).append(e.duration_ns() / 1e3)
    for k, v in agg.items():
        print("%-40s n=%2d  %s" % (k, len(v), " ".join("%.1f" % x for x in v[:6])))
    print("bw", bw[:4].tolist())


if __name__ == "__main__":
    main()
