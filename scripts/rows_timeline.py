#!/usr/bin/env python
"""Phase timeline of the tensor-core K-seed kernels (one CTA), from the in-kernel clock64 marks."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import _lib, ops, pipeline, synthetic  # noqa: E402

NAMES = {1: "iter start", 2: "prep done", 3: "operand tile written", 4: "S ready", 5: "coeff computed", 6: "coeff tile free",
         7: "coeff tile written", 8: "flush prev done", 9: "flush last done", 10: "dY ready", 11: "partials written",
         12: "cluster sync 1", 13: "reduce done", 14: "cluster sync 2"}


def dump(tag):
    lib = _lib.load()
    buf = (ctypes.c_longlong * 8192)()
    lib.prifit_debug_rows_timeline.restype = ctypes.c_int
    n = lib.prifit_debug_rows_timeline(buf, 4096)
    print("== %s: %d marks" % (tag, n))
    prev = None
    agg = {}
    for i in range(n):
        pid, clk = buf[2 * i], buf[2 * i + 1]
        if prev is not None:
            key = (prev[0], pid)
            agg.setdefault(key, []).append(clk - prev[1])
        prev = (pid, clk)
    tot = sum(sum(v) for v in agg.values())
    for (p0, p1), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("  %-22s -> %-22s n=%3d  mean %7.0f clk  total %5.1f%%" % (NAMES.get(p0, p0), NAMES.get(p1, p1), len(v), sum(v) / len(v), 100.0 * sum(v) / tot))
    print("  total %.1f us at 1.965 GHz" % (tot / 1965.0))


def main():
    B, N = int(os.environ.get("B", 24)), int(os.environ.get("N", 2048))
    dev = torch.device("cuda:0")
    E, _, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=0)
    X = ops.normalize_fwd(E.to(dev))
    res = pipeline.cluster_batch(X, N, 0.05, 10, 25)
    gC = torch.randn(B, res.kcap, 128, device=dev)
    gX = torch.zeros_like(X)
    os.environ["PRIFIT_ROWS_TIMELINE"] = os.environ.get("ROWS_DBG", "1")
    traj, stat, C = ops.rows_fwd(X, res.bw, res.idx, res.K, 10, res.kcap, 0)
    torch.cuda.synchronize()
    dump("forward")
    ops.rows_bwd(X, res.bw, res.idx, res.K, traj, stat, gC, gX, 10, res.kcap, 0)
    torch.cuda.synchronize()
    dump("backward")


if __name__ == "__main__":
    main()
