#!/usr/bin/env python
"""cProfile of the host side of cfg2 steps (where the CPU spends its time between launches)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import pipeline, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    E, P, _ = synthetic.planted_shapes(24, n_points=2048, n_clusters=16, seed=0)
    E, P = E.to(dev), P.to(dev)

    cf = "--cf" in sys.argv          # the reference-shaped API on channel-first tensors (what bench.py's e2e arm calls)
    if cf:
        import prifit_b200.convex_loss as cl
        Ecf, Pcf = E.permute(0, 2, 1).contiguous(), P.permute(0, 2, 1).contiguous()

    def step():
        if cf:
            Ei = Ecf.detach().requires_grad_(True)
            total, l, params, labels = cl.convex_loss(Pcf, Pcf, Ei, quantile=0.05, iterations=10, max_num_clusters=25, full_chamfer=False)
            total.backward()
            return
        Ei = E.detach().requires_grad_(True)
        out = pipeline.fit_loss(Ei, P, quantile=0.05, iterations=10, max_num_clusters=25)
        out["loss"].backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(40)


if __name__ == "__main__":
    main()
