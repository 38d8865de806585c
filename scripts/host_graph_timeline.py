#!/usr/bin/env python
"""Host-side marks of a graph-replayed step: how long after the guard read-back (end of graph 1) does the host need
to get graph 3 launched?  The device only has graph 2 (~0.2 ms) queued during that time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import graph_step, pipeline, synthetic  # noqa: E402
import prifit_b200.convex_loss as cl  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cf = "--cf" in sys.argv
    E, P, _ = synthetic.planted_shapes(24, n_points=2048, n_clusters=16, seed=0)
    E, P = E.to(dev), P.to(dev)
    if cf:
        E, P = E.permute(0, 2, 1).contiguous(), P.permute(0, 2, 1).contiguous()
    marks = []
    orig_sync = torch.cuda.Event.synchronize
    orig_replay = torch.cuda.CUDAGraph.replay

    def sync(self):
        marks.append(("sync enter", time.perf_counter()))
        r = orig_sync(self)
        marks.append(("sync exit", time.perf_counter()))
        return r

    def replay(self):
        r = orig_replay(self)
        marks.append(("graph launched", time.perf_counter()))
        return r

    torch.cuda.Event.synchronize = sync
    torch.cuda.CUDAGraph.replay = replay
    for it in range(8):
        marks.clear()
        t0 = time.perf_counter()
        Ei = E.detach().requires_grad_(True)
        if cf:
            total, l, params, labels = cl.convex_loss(P, P, Ei, quantile=0.05, iterations=10, max_num_clusters=25, full_chamfer=False)
        else:
            total = pipeline.fit_loss(Ei, P, quantile=0.05, iterations=10, max_num_clusters=25)["loss"]
        marks.append(("forward returned", time.perf_counter()))
        total.backward()
        marks.append(("backward returned", time.perf_counter()))
        if it >= 5:
            print("step %d (%s)" % (it, "convex_loss" if cf else "fit_loss"))
            for name, t in marks:
                print("   %-20s %8.1f us" % (name, (t - t0) * 1e6))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
