#!/usr/bin/env python
"""Device timeline of steady-state cfg2 steps (CUPTI through torch.profiler): every kernel / memcpy with
its start offset, duration and the idle gap in front of it -- shows where the GPU waits for the host."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import dist as pdist, pipeline, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, N = 24, 2048
    sets = []
    for s in range(4):
        E, P, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=1000 + s * B)
        sets.append((E.to(dev), P.to(dev)))

    def step(i):
        E, P = sets[i % 4]
        Ei = E.detach().requires_grad_(True)
        out = pipeline.fit_loss(Ei, P, quantile=0.05, iterations=10, max_num_clusters=25)
        L, Lb = pdist.global_loss(out)
        Lb.backward()

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    n_steps = 6
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(n_steps):
            step(i)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # steps start at normalize_fwd_kernel
    starts = [i for i, e in enumerate(evs) if "normalize_fwd" in e.name]
    if len(starts) < 4:
        print("could not find step boundaries (%d kernels traced)" % len(evs))
        return
    lo, hi = starts[2], starts[3]
    t0 = evs[lo].time_range.start
    prev_end = evs[lo - 1].time_range.end if lo > 0 else t0
    busy = 0.0
    print("%9s %8s %8s  %s" % ("start us", "dur us", "gap us", "kernel"))
    for e in evs[lo:hi]:
        s, d = e.time_range.start, e.time_range.end - e.time_range.start
        gap = s - prev_end
        busy += d
        print("%9.1f %8.1f %8.1f  %s" % (s - t0, d, gap, e.name[:90]))
        prev_end = max(prev_end, e.time_range.end)
    span = evs[hi].time_range.start - t0
    print("step span %.1f us, busy %.1f us, idle %.1f us, %d device activities" % (span, busy, span - busy, hi - lo))
    for k in range(1, len(starts) - 1):
        print("step %d span %.1f us" % (k, evs[starts[k + 1]].time_range.start - evs[starts[k]].time_range.start))


if __name__ == "__main__":
    main()
