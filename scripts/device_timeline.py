#!/usr/bin/env python
"""Device timeline of steady-state cfg2 steps (CUPTI through torch.profiler): every kernel / memcpy with
its start offset, duration and the idle gap in front of it -- shows where the GPU waits for the host."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import dist as pdist, pipeline, synthetic  # noqa: E402


def print_timeline(prof, out=sys.stdout, first_kernel="normalize_fwd", last_kernel="normalize_bwd", which=2):
    """Prints the device activities of step number `which` (a step starts at the first `first_kernel` that follows a
    `last_kernel`) with start offset, duration and stream, plus busy / idle totals."""
    p = lambda *a: print(*a, file=out)
    kin = prof.profiler.kineto_results.events()
    evs = []
    for e in kin:
        if e.device_type() == torch.autograd.DeviceType.CUDA:
            evs.append((e.start_ns() / 1e3, e.duration_ns() / 1e3, e.name(), e.device_resource_id()))
    evs.sort()
    # a step starts at the first normalize_fwd that follows a normalize_bwd
    starts, seen_bwd = [], True
    for i, e in enumerate(evs):
        if last_kernel in e[2]:
            seen_bwd = True
        elif first_kernel in e[2] and seen_bwd:
            starts.append(i)
            seen_bwd = False
    if len(starts) < 4:
        p("could not find step boundaries (%d device activities traced)" % len(evs))
        return
    lo, hi = starts[which], starts[which + 1]
    t0 = evs[lo][0]
    streams = sorted({e[3] for e in evs[lo:hi]})
    p("%9s %8s %3s  %s" % ("start us", "dur us", "st", "kernel"))
    # union of busy intervals = time with at least one kernel resident
    busy, cur_s, cur_e, ksum = 0.0, None, None, 0.0
    for s_, d_, name, st in evs[lo:hi]:
        p("%9.1f %8.1f %3d  %s" % (s_ - t0, d_, streams.index(st), name[:90]))
        ksum += d_
        if cur_e is None or s_ > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s_, s_ + d_
        else:
            cur_e = max(cur_e, s_ + d_)
    busy += cur_e - cur_s
    span = evs[hi][0] - t0
    p("step span %.1f us, some kernel resident %.1f us, idle %.1f us, sum of kernel durations %.1f us, %d activities on %d streams"
          % (span, busy, span - busy, ksum, hi - lo, len(streams)))
    for k in range(1, len(starts) - 1):
        p("step %d span %.1f us" % (k, evs[starts[k + 1]][0] - evs[starts[k]][0]))




def main():
    dev = torch.device("cuda:0")
    B, N = 24, 2048
    sets = []
    for s in range(4):
        E, P, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=1000 + s * B)
        sets.append((E.to(dev), P.to(dev)))

    cf = "--cf" in sys.argv            # the reference-shaped public API on channel-first tensors (convex_loss)
    if cf:
        import prifit_b200.convex_loss as cl
        sets = [(E.permute(0, 2, 1).contiguous(), P.permute(0, 2, 1).contiguous()) for E, P in sets]

    def step(i):
        E, P = sets[i % 4]
        Ei = E.detach().requires_grad_(True)
        if cf:
            total, l, params, labels = cl.convex_loss(P, P, Ei, quantile=0.05, iterations=10, max_num_clusters=25, full_chamfer=False)
            total.backward()
            return
        out = pipeline.fit_loss(Ei, P, quantile=0.05, iterations=10, max_num_clusters=25, graph="--eager" not in sys.argv)
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
    print_timeline(prof)


if __name__ == "__main__":
    main()
