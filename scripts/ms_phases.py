#!/usr/bin/env python
"""In-kernel phase clocks of the all-seed kernel (PRIFIT_MS_PAIR=0 PRIFIT_MS_DBG=3): one launch, CTA (0,0) prints."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import ops, synthetic  # noqa: E402
dev = torch.device("cuda:0")
for B in (24, 1):
    E, _, _ = synthetic.planted_shapes(B, n_points=2048, n_clusters=16, seed=0)
    X = ops.normalize_fwd(E.to(dev))
    bw = torch.full((B,), 0.15, device=dev)
    print("B =", B, flush=True)
    ops.meanshift(X, bw, 10, ops.MS_F16_TCGEN05)
    torch.cuda.synchronize()
