"""GPU: the tensor-core bandwidth path really is the path taken (no silent fallback to the CUDA-core
kernel) on ordinary inputs, and the on-device fallback fires only when a candidate list overflows."""
import numpy as np
import pytest
import torch

from oracle import restatement as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,q", [(2048, 0.05), (10000, 0.05), (777, 0.2)])
def test_tensor_core_bandwidth_does_not_fall_back(cuda, n, q):
    from prifit_b200 import ops, synthetic

    E, _, _ = synthetic.planted_shapes(3, n_points=n, n_clusters=8, seed=17)
    X = R.normalize_twice(E).to(cuda)
    k = torch.full((3,), int(q * n), dtype=torch.int32, device=cuda)
    bw = ops.bandwidth(X, k)
    assert not ops.last_bandwidth_fell_back()
    ref = float(R.compute_bandwidth(X[1].cpu(), n, q, perm=np.arange(n)))
    assert abs(float(bw[1]) - ref) / ref < 2e-6


def test_duplicates_trigger_the_device_fallback(cuda):
    from prifit_b200 import ops

    g = torch.Generator().manual_seed(3)
    x = torch.nn.functional.normalize(torch.randn(1, 128, generator=g), dim=1)
    X = x.repeat(600, 1)[None].to(cuda)
    bw = ops.bandwidth(X, torch.tensor([30], dtype=torch.int32, device=cuda))
    assert ops.last_bandwidth_fell_back()
    assert float(bw[0]) == pytest.approx(1e-3, rel=1e-4)
