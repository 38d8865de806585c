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


@pytest.mark.parametrize("family", ["hier", "smooth", "unbalanced", "random"])
def test_log_binned_bandwidth_is_exact_over_distance_scales(cuda, family):
    """The logarithmic histogram level covers distances from 6e-5 to 4: tight clusters (hier: k-th distances ~ 0.006),
    smooth embeddings without modes, unbalanced noisy clusters, i.i.d. directions (distances ~ 2) -- the tensor-core path
    must return the exact fp32 order statistic of the CUDA-core kernel for several quantiles, without falling back."""
    from prifit_b200 import _lib, ops, synthetic

    if family == "hier":
        E, _, _ = synthetic.hier_shapes(2, seed=11)
    elif family == "smooth":
        E, _ = synthetic.smooth_shapes(2, n_points=1536, freq=0.5, seed=12)
    elif family == "unbalanced":
        E, _, _ = synthetic.unbalanced_shapes(2, n_points=2048, sigma=0.03, seed=13)
    else:
        E, _ = synthetic.random_shapes(2, n_points=1280, seed=14)
    X = R.normalize_twice(E).to(cuda)
    n = X.shape[1]
    for q in (0.01, 0.05, 0.3):
        k = torch.full((2,), int(q * n), dtype=torch.int32, device=cuda)
        bw = ops.bandwidth(X, k)
        assert not ops.last_bandwidth_fell_back(), (family, q)
        prev = _lib.load().prifit_set_gram_engine(1)
        try:
            exact = ops.bandwidth(X, k)
        finally:
            _lib.load().prifit_set_gram_engine(prev)
        # both are exact fp32 order statistics; they differ by the summation order of the 128-term dot product behind
        # 2 - 2 <x_i, x_j> (one ulp of a number near 1 = 6e-8 on a distance that can be as small as 6e-3: hier, q = 0.01)
        assert float((bw - exact).abs().max() / exact.abs().max()) < 4e-6, (family, q, bw.tolist(), exact.tolist())
