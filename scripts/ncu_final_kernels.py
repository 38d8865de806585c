#!/usr/bin/env python
"""The tensor-core kernels of the cluster stage at cfg2, one launch each, for a single `ncu --set full` capture:
bandwidth (gram_tc_kernel<HIST> level 0, the level-1 launch that exits, <COLLECT>), the all-seed mean-shift kernel, the NMS
arg-min pass (gram_tc_kernel<NEAREST>).  No warm-up: ncu replays every captured launch ~40 times with its own cache control.

    ncu --set full --clock-control none --import-source on -k regex:'meanshift_tc_kernel|gram_tc_kernel' -c 6 -f \
        -o gpurun_out/prof_r02_final_kernels python scripts/ncu_final_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prifit_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
B, N = 24, 2048
E, _, _ = synthetic.planted_shapes(B, n_points=N, n_clusters=16, seed=0)
X = ops.normalize_fwd(E.to(dev))
bw = ops.bandwidth(X, torch.full((B,), int(0.05 * N), dtype=torch.int32, device=dev))
newX = ops.meanshift(X, bw, 10, ops.MS_F16_TCGEN05)
idx, K, labels, nlab = ops.nms(newX, bw, 32)
torch.cuda.synchronize()
print("bw", bw[:3].tolist(), "K", K[:6].tolist())
