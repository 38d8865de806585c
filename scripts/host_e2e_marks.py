#!/usr/bin/env python
"""Host-side marks of the end-to-end loop (convex_loss on channel-first tensors staged from pinned memory, backward, loss
read-back): where the host spends the step, in particular between the arrival of the guard inputs and the next step's
graph launch -- the window in which the device only has the latency chains (~0.5 ms) queued."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import graph_step, synthetic  # noqa: E402
import prifit_b200.convex_loss as cl  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    E, P, _ = synthetic.planted_shapes(24, n_points=2048, n_clusters=16, seed=0)
    hX = [E.permute(0, 2, 1).contiguous().pin_memory() for _ in range(2)]
    hP = [P.permute(0, 2, 1).contiguous().pin_memory() for _ in range(2)]
    marks = []
    mark = lambda name: marks.append((name, time.perf_counter()))
    GS = graph_step.GraphStep
    for name in ("run_forward", "finish_forward", "run_backward"):
        orig = getattr(GS, name)

        def wrap(self, *a, _o=orig, _n=name, **k):
            mark(_n + " >")
            r = _o(self, *a, **k)
            mark(_n + " <")
            return r
        setattr(GS, name, wrap)
    orig_replay = torch.cuda.CUDAGraph.replay

    def replay(self):
        r = orig_replay(self)
        mark("graph launched")
        return r
    torch.cuda.CUDAGraph.replay = replay
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}

    def stage(i):
        with torch.cuda.stream(copy_stream):
            X = hX[i % 2].to(dev, non_blocking=True)
            pts = hP[i % 2].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[i] = (X, pts, ev)
        mark("staged next")

    stage(0)
    t_prev = None
    for it in range(10):
        marks.clear()
        t0 = time.perf_counter()
        X, pts, ev = staged.pop(it)
        torch.cuda.current_stream().wait_event(ev)
        X = X.requires_grad_(True)
        graph_step.enqueued_hook = lambda: stage(it + 1)
        total, l, params, labels = cl.convex_loss(pts, pts, X, quantile=0.05, iterations=10, max_num_clusters=25, full_chamfer=False)
        graph_step.enqueued_hook = None
        mark("convex_loss returned")
        total.backward()
        mark("backward returned")
        if it >= 7:
            print("step %d   (previous step's loop took %.1f us)" % (it, 0.0 if t_prev is None else (t0 - t_prev) * 1e6))
            for name, t in marks:
                print("   %-22s %8.1f us" % (name, (t - t0) * 1e6))
        t_prev = t0
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
