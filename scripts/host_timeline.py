#!/usr/bin/env python
"""Host vs device time of one cfg2 step: where does the GPU wait for the CPU?"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import ops, pipeline, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, N = 24, 2048
    E, P, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=0)
    E, P = E.to(dev), P.to(dev)
    marks = []
    orig_cluster = pipeline.cluster_batch

    def cluster_batch(*a, **k):
        marks.append(("cluster enter", time.perf_counter()))
        r = orig_cluster(*a, **k)
        marks.append(("cluster exit (counts on host)", time.perf_counter()))
        return r

    pipeline.cluster_batch = cluster_batch
    for it in range(6):
        marks.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record()
        Ei = E.detach().requires_grad_(True)
        out = pipeline.fit_loss(Ei, P, quantile=0.05, iterations=10, max_num_clusters=25)
        t_fwd = time.perf_counter()
        e1.record()
        out["loss"].backward()
        t_bwd = time.perf_counter()
        e2.record()
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        if it >= 3:
            print("host: fwd enqueued %.2f ms, bwd enqueued %.2f ms, all done %.2f ms | device: fwd %.2f ms, bwd %.2f ms" % (
                (t_fwd - t0) * 1e3, (t_bwd - t0) * 1e3, (t_end - t0) * 1e3, e0.elapsed_time(e1), e1.elapsed_time(e2)))
            for name, t in marks:
                print("   %-32s at %.2f ms" % (name, (t - t0) * 1e3))


if __name__ == "__main__":
    main()
