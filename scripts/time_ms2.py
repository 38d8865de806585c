#!/usr/bin/env python
"""All-seed kernel timing under the conditions of bench.py's eager step: real bandwidths, one launch between other work."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import ops, synthetic  # noqa: E402
from scripts.time_ms import timeit  # noqa: E402
dev = torch.device("cuda:0")
B, N = 24, 2048
E, _, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=0)
X = ops.normalize_fwd(E.to(dev))
flops = 4.0 * N * N * 128 * 10 * B
for bwv in (0.15, 0.05, 0.02):
    bw = torch.full((B,), bwv, device=dev)
    t = timeit(lambda: ops.meanshift(X, bw, 10, ops.MS_F16_TCGEN05))
    print("bw=%.2f back-to-back  %8.1f us  %.0f TFLOP/s" % (bwv, t, flops / t / 1e6), flush=True)
bw = ops.bandwidth(X, torch.full((B,), int(0.05 * N), dtype=torch.int32, device=dev))
print("real bw", bw[:4].tolist())
ts = []
big = torch.empty(64 << 20, device=dev)
for i in range(12):
    big.normal_()                    # other work + L2 flush between launches
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.meanshift(X, bw, 10, ops.MS_F16_TCGEN05); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("isolated launches (us):", " ".join("%.0f" % t for t in ts))
